"""Loss assembly of the reference's workers, restricted to the hot path (photometric + smoothness
+ auxiliary L1 terms); mirrors

    single_frame_worker.Worker.loss_forward   reference model/single_frame_worker.py:101-165
    multi_frame_worker.Worker.loss_forward    reference model/multi_frame_worker.py:103-175

Same weights and the same list-of-0-dim-tensors return convention (the worker sums it,
model/worker.py:522).  The geometric (flow-consistency) terms are appended by the caller.
"""
import torch

from .networks import DisparitySmoothLoss, RectifiedPatternSimilarityLoss


def _merge(x):
    """[tl,bs,C,H,W] -> [tl*bs,C,H,W] (model/multi_frame_networks.py:36-37); 4-D tensors pass through."""
    return x.contiguous().view(-1, *x.shape[-3:]) if x.dim() == 5 else x


class _HotPathLoss(torch.nn.Module):
    smooth_weight = None

    def __init__(self, im_height, im_width, pattern, loss_type='census_sad', loss_eps=0.5, block_size=9,
                 process_group=None):
        super().__init__()
        self.ph_loss = RectifiedPatternSimilarityLoss(im_height, im_width, pattern, loss_type, loss_eps,
                                                      block_size=block_size, return_pattern_proj=False,
                                                      process_group=process_group)
        self.disparity_loss = DisparitySmoothLoss(process_group=process_group)


class SingleFrameLoss(_HotPathLoss):
    """out: list of per-scale disparities (all full resolution, model/networks.py:290-295)."""

    def forward(self, out, im_lcn, std, ambient, pseudo_gt=None):
        if not isinstance(out, (tuple, list)):
            out = [out]
        im, std, amb = _merge(im_lcn)[:, 0:1], _merge(std), _merge(ambient)
        ph = self.ph_loss.forward_multi([_merge(o) for o in out], im, std)   # :108-115, all scales fused
        vals = [v / (2 ** s) for s, v in enumerate(ph)]
        vals.append(self.disparity_loss(_merge(out[0]), amb) * 0.4)   # :118-124 (scale 0 only)
        if pseudo_gt is not None:                                     # :152-155 (DIS-FTSF)
            for s, o in enumerate(out):
                vals.append(torch.mean(torch.abs(o - pseudo_gt)) * 0.1 / (2 ** s))
        return vals


class MultiFrameLoss(_HotPathLoss):
    def forward(self, out, im_lcn, std, ambient, primary_disp=None):
        if not isinstance(out, (tuple, list)):
            out = [out]
        im, std, amb = _merge(im_lcn)[:, 0:1], _merge(std), _merge(ambient)
        ph = self.ph_loss.forward_multi([_merge(o) for o in out], im, std)   # :110-117
        vals = [v / (2 ** s) for s, v in enumerate(ph)]
        vals.append(self.disparity_loss(_merge(out[0]), amb) * 0.8)   # :120-126
        if primary_disp is not None:                                  # :160-165 (first two epochs)
            vals.append(torch.mean(torch.abs(out[0] - primary_disp)) * 0.1)
        return vals
