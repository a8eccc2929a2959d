"""Host-side mirror of the hot-path modules of the reference's model/networks.py.

Same constructor / call signatures and return values as the reference classes, with every
forward and backward executed by libdis_b200.so:

    LCN(radius, epsilon)(data) -> (lcn, std)                                   reference :663-689
    RectifiedPatternSimilarityLoss(im_height, im_width, pattern,
        loss_type='census_sad', loss_eps=0.5)(disp0, im, std=None,
        output_mean=True) -> (val, pattern_proj)                               reference :336-377
    DisparitySmoothLoss()(disp, im) -> scalar                                  reference :411-431
    SobelFilter(norm=False, ksize=5)(x) -> [N,2,H,W] | [N,1,H,W]               reference :693-730

Differences that are deliberate:
  * no torch.cuda.synchronize() around forward (the reference's TimedModule does, :66-71);
  * batch-wide ratios can be formed over a torch.distributed process group
    (``process_group=``) so that data-parallel ranks reproduce single-process semantics;
  * ``RectifiedPatternSimilarityLoss(..., return_pattern_proj=False)`` skips materialising
    pattern_proj (the workers ignore it, single_frame_worker.py:114).
"""
import torch

from . import _ops
from .parallel import all_reduce_sum_


class LCN(torch.nn.Module):
    """Local contrast normalisation, reference model/networks.py:663-689."""

    def __init__(self, radius, epsilon):
        super().__init__()
        self.radius = radius
        self.epsilon = epsilon

    def forward(self, data):
        if data.requires_grad and torch.is_grad_enabled():
            return _LCNFunction.apply(data, self.radius, self.epsilon)
        return _ops.lcn_forward(data, self.radius, self.epsilon)

    tforward = forward

    def prepare_input(self, im):
        """Worker.copy_data for an image key (reference model/worker.py:418-438), fused into one launch:
        im [bs,tl,1,H,W] as delivered by the DataLoader -> (im_cat [tl,bs,2,H,W] = cat(LCN(im), im), std [tl,bs,1,H,W])."""
        return _ops.lcn_prepare_input(im, self.radius, self.epsilon)


class _LCNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, radius, eps):
        ctx.set_materialize_grads(False)
        lcn, std = _ops.lcn_forward(data, radius, eps)
        ctx.save_for_backward(data, lcn, std)
        ctx.radius, ctx.eps = radius, eps
        return lcn, std

    @staticmethod
    def backward(ctx, g_lcn, g_std):
        data, lcn, std = ctx.saved_tensors
        return _ops.lcn_backward(data, lcn, std, g_lcn, g_std, ctx.radius, ctx.eps), None, None


class _PatternLossMean(torch.autograd.Function):
    """val = sum(mask * diff) / sum(mask) with the un-normalised gradient stashed by the forward pass."""

    @staticmethod
    def forward(ctx, disp0, im, std, pattern, block_size, type_id, eps, want_proj, group):
        ctx.set_materialize_grads(False)
        need_grad = ctx.needs_input_grad[0]
        out3, proj, _, gnum = _ops.pattern_loss_forward(disp0, im, std, pattern, block_size, type_id, eps,
                                                        want_proj=want_proj, want_diff=False, want_grad=need_grad)
        if group is not None:
            all_reduce_sum_(out3[:2], group)
            val = out3[0] / out3[1]
        else:
            val = out3[2].clone()
        ctx.save_for_backward(gnum, out3, disp0 if want_proj else None, pattern if want_proj else None)
        if proj is None:
            proj = disp0.new_empty(0)
            ctx.mark_non_differentiable(proj)
        return val, proj

    @staticmethod
    def backward(ctx, g_val, g_proj):
        gnum, out3, disp0, pattern = ctx.saved_tensors
        if g_val is None:
            g_val = torch.zeros((), dtype=torch.float32, device=gnum.device)
        grad = _ops.scale_by_device_scalar(gnum, g_val, out3[1:2])
        if g_proj is not None and disp0 is not None and g_proj.numel() == disp0.numel():
            _, dproj, _, _ = _ops.pattern_warp(disp0, pattern, want_dproj=True)
            grad = grad + _ops.mul(g_proj.contiguous(), dproj)
        return grad, None, None, None, None, None, None, None, None


class _PatternLossMulti(torch.autograd.Function):
    """S (2 or 4) scales of the same frames in one launch: returns the S ratios."""

    @staticmethod
    def forward(ctx, im, std, pattern, block_size, type_id, eps, group, *disps):
        ctx.set_materialize_grads(False)
        S = len(disps)
        need = [ctx.needs_input_grad[7 + i] for i in range(S)]
        out3, grads = _ops.pattern_loss_multi_forward(list(disps), im, std, pattern, block_size, type_id, eps,
                                                      want_grad=any(need))
        if group is not None:
            nd = out3[:, :2].contiguous()
            all_reduce_sum_(nd, group)
            out3 = torch.cat((nd, (nd[:, 0] / nd[:, 1]).unsqueeze(1)), dim=1)
        ctx.S = S
        ctx.save_for_backward(out3, *(grads if grads is not None else []))
        return tuple(out3[i, 2].clone() for i in range(S))

    @staticmethod
    def backward(ctx, *g_vals):
        out3, *grads = ctx.saved_tensors
        res = []
        for i in range(ctx.S):
            if g_vals[i] is None or not grads or not ctx.needs_input_grad[7 + i]:
                res.append(None)
            else:
                res.append(_ops.scale_by_device_scalar(grads[i], g_vals[i], out3[i, 1:2]))
        return (None,) * 7 + tuple(res)


class _PatternLossMap(torch.autograd.Function):
    """output_mean=False: per-pixel loss map (reference :375-376)."""

    @staticmethod
    def forward(ctx, disp0, im, std, pattern, block_size, type_id, eps):
        ctx.set_materialize_grads(False)
        _, proj, diff, _ = _ops.pattern_loss_forward(disp0, im, None, pattern, block_size, type_id, eps,
                                                     want_proj=True, want_diff=True, want_grad=False)
        ctx.save_for_backward(disp0, im, pattern, proj)
        ctx.cfg = (block_size, type_id, eps)
        return diff, proj

    @staticmethod
    def backward(ctx, g_diff, g_proj):
        disp0, im, pattern, proj = ctx.saved_tensors
        block_size, type_id, eps = ctx.cfg
        if g_diff is not None:
            g_e = _ops.photometric_loss_backward(proj, im, g_diff.contiguous(), block_size, type_id, eps)
            if g_proj is not None:
                g_e = g_e + g_proj
        elif g_proj is not None:
            g_e = g_proj.contiguous()
        else:
            return (None,) * 7
        _, dproj, _, _ = _ops.pattern_warp(disp0, pattern, want_dproj=True)
        return _ops.mul(g_e, dproj), None, None, None, None, None, None


class RectifiedPatternSimilarityLoss(torch.nn.Module):
    """Photometric loss between the disparity-warped projector pattern and the LCN image,
    reference model/networks.py:336-377."""

    def __init__(self, im_height, im_width, pattern, loss_type='census_sad', loss_eps=0.5,
                 block_size=9, return_pattern_proj=True, process_group=None):
        super().__init__()
        self.im_height = im_height
        self.im_width = im_width
        # the reference averages the (3 identical) channels once, :344
        self.pattern = pattern.mean(dim=1, keepdim=True).contiguous()
        self.loss_type = loss_type
        self.loss_eps = loss_eps
        self.block_size = block_size          # hard-coded 9 in the reference, :372
        self.return_pattern_proj = return_pattern_proj
        self.process_group = process_group

    def forward(self, disp0, im, std=None, output_mean=True):
        if tuple(disp0.shape[-2:]) != (self.im_height, self.im_width):
            raise ValueError(f"disp0 is {tuple(disp0.shape)}, loss was built for {self.im_height}x{self.im_width}")
        self.pattern = self.pattern.to(device=disp0.device, dtype=torch.float32)
        type_id = _ops.loss_type_id(self.loss_type)
        im = im.contiguous()
        if output_mean:
            val, proj = _PatternLossMean.apply(disp0, im, std, self.pattern, self.block_size, type_id, self.loss_eps,
                                               self.return_pattern_proj, self.process_group)
            return val, (proj if self.return_pattern_proj else None)
        return _PatternLossMap.apply(disp0, im, std, self.pattern, self.block_size, type_id, self.loss_eps)

    tforward = forward

    def forward_multi(self, disps, im, std=None):
        """The worker's photometric loop (model/single_frame_worker.py:108-115) in as few launches as possible:
        groups of 4 / 2 scales go through one launch (census types: the packed multi-scale kernel; mse / sad: the
        point-wise kernel behind one box filter of the weights), the rest one by one.
        -> list of 0-dim ratios, one per disparity map (un-weighted)."""
        self.pattern = self.pattern.to(device=im.device, dtype=torch.float32)
        type_id = _ops.loss_type_id(self.loss_type)
        im = im.contiguous()
        vals, i = [], 0
        while i < len(disps):
            left = len(disps) - i
            take = 4 if left >= 4 else (2 if left >= 2 else 1)
            if take == 1:
                v, _ = _PatternLossMean.apply(disps[i], im, std, self.pattern, self.block_size, type_id, self.loss_eps,
                                              False, self.process_group)
                vals.append(v)
                take = 1
            else:
                vals.extend(_PatternLossMulti.apply(im, std, self.pattern, self.block_size, type_id, self.loss_eps,
                                                    self.process_group, *disps[i:i + take]))
            i += take
        return vals


class _SobelFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ksize):
        ctx.ksize = ksize
        return _ops.sobel_forward(x, ksize)

    @staticmethod
    def backward(ctx, g):
        return _ops.sobel_backward(g.contiguous(), ctx.ksize), None


class SobelFilter(torch.nn.Module):
    """reference model/networks.py:693-730 (weights are constants here, not nn.Parameters: the reference
    never optimises them and only pays for their useless weight gradients)."""

    def __init__(self, norm=False, ksize=5):
        super().__init__()
        if ksize not in (3, 5):
            raise ValueError("ksize must be 3 or 5")
        self.ksize = ksize
        self.norm = norm

    def forward(self, x):
        g = _SobelFunction.apply(x, self.ksize)
        if self.norm:
            return torch.sqrt(g[:, 0:1] ** 2 + g[:, 1:2] ** 2 + 1e-8)
        return g

    tforward = forward


class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, im, group, fork, side):
        # fork = (event, buffers), side = stream (optional).  The kernel is then launched on `side` once the event has passed,
        # so that it shares the GPU with work queued on the current stream after the event (losses._HotPathLoss: the tail
        # wave of the photometric kernel).  Its outputs were allocated BEFORE the event was recorded (memory handed out
        # later could still be in use by that queued work); the streams are joined right here, so every later reuse of the
        # tensors involved is ordered behind the side stream's work.  The autograd node belongs to the current stream.
        want = ctx.needs_input_grad[0]
        if side is None or (want and fork[1][1] is None):
            out3, gsum = _ops.smooth_loss_forward(disp, im, want_grad=want)
        else:
            main = torch.cuda.current_stream(disp.device)
            side.wait_event(fork[0])
            out3, gsum = _ops.smooth_loss_forward(disp, im, want_grad=want, launch_stream=side, buffers=fork[1])
            main.wait_stream(side)
        if group is not None:
            all_reduce_sum_(out3[:2], group)
            val = out3[0] / out3[1]
        else:
            val = out3[2].clone()
        ctx.save_for_backward(gsum, out3)
        return val

    @staticmethod
    def backward(ctx, g_val):
        gsum, out3 = ctx.saved_tensors
        return _ops.scale_by_device_scalar(gsum, g_val, out3[1:2]), None, None, None, None


class DisparitySmoothLoss(torch.nn.Module):
    """Edge-aware smoothness, reference model/networks.py:411-431 (gradient flows to disp only:
    the ambient image is data)."""

    def __init__(self, process_group=None):
        super().__init__()
        self.process_group = process_group

    def forward(self, disp, im, fork=None, side_stream=None):
        return _SmoothLoss.apply(disp, im.contiguous(), self.process_group, fork, side_stream)

    tforward = forward


# ----------------------------------------------------------------------------------------------------
# Flow-consistency (geometric) losses, reference model/networks.py:433-661
# ----------------------------------------------------------------------------------------------------
class _FlowConsistency(torch.autograd.Function):
    """Both directions (0->1 and 1->0) of the loss: two fused launches; gradients w.r.t. depth0 and depth1."""

    @staticmethod
    def forward(ctx, depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, primary0, primary1, K, ray, clamp,
                multi_frame, group=None):
        ctx.set_materialize_grads(False)
        need0, need1 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        # direction A: frame 0 -> 1 (direct gradient to depth0, scatter to depth1); direction B: the mirror image
        a3, mask0, orig, gA0, gA1 = _ops.flow_consistency_dir(depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, K,
                                                              ray, clamp, primary1 if multi_frame else None,
                                                              True, not multi_frame, need0, need1)
        b3, mask1, _, gB1, gB0 = _ops.flow_consistency_dir(depth1, depth0, R1, t1, R0, t0, flow1, flow0, amb1, amb0, K,
                                                           ray, clamp, primary0 if multi_frame else None,
                                                           True, False, need1, need0)
        if group is not None:      # batch-wide ratios: all-reduce both (numerator, denominator) pairs in one vector
            nd = torch.stack((a3[:2], b3[:2]))
            all_reduce_sum_(nd, group)
            a3, b3 = nd[0], nd[1]
        loss = a3[0] / (a3[1] + 1e-8) + b3[0] / (b3[1] + 1e-8)      # (diff*mask).sum() / (mask.sum() + 1e-8), :599, :653
        ctx.save_for_backward(a3, b3, gA0, gA1, gB0, gB1)
        if orig is None:
            orig = depth0.new_empty(0)
        ctx.mark_non_differentiable(mask0, mask1, orig)
        return loss, mask0, mask1, orig

    @staticmethod
    def backward(ctx, g_loss, *_):
        a3, b3, gA0, gA1, gB0, gB1 = ctx.saved_tensors
        if g_loss is None:
            return (None,) * 17
        g0 = _ops.combine2(gA0, gB0, g_loss, a3[1:2], b3[1:2], 1e-8) if gA0 is not None else None
        g1 = _ops.combine2(gB1, gA1, g_loss, b3[1:2], a3[1:2], 1e-8) if gB1 is not None else None
        return (g0, g1) + (None,) * 15


class ProjectionBaseLoss(torch.nn.Module):
    """Holds K and the per-pixel rays [u, v, 1] @ Ki^T (computed in float64, stored float32 like the reference,
    model/networks.py:437-453)."""

    def __init__(self, K, Ki, im_height, im_width, process_group=None):
        super().__init__()
        import numpy as np
        self.process_group = process_group
        self.K = K.reshape(3, 3).to(torch.float32)
        self.im_height, self.im_width = im_height, im_width
        u, v = np.meshgrid(range(im_width), range(im_height))
        uv = np.stack((u, v, np.ones_like(u)), axis=2).reshape(-1, 3)
        self.ray = torch.from_numpy((uv @ Ki.detach().cpu().numpy().T).astype(np.float32))

    def _consts(self, ref):
        self.K = self.K.to(ref.device)
        self.ray = self.ray.to(ref.device)
        return self.K, self.ray


class Single_Frame_Flow_Consistency_Loss(ProjectionBaseLoss):
    """reference model/networks.py:609-661.  Returns (loss, mask0, mask1, orig_mask) like the reference.  By default
    orig_mask ([H,W], first sample) stays a device tensor: the reference's blocking `.to('cpu').numpy()` (:640)
    stalls the stream for a value the worker never uses (single_frame_worker.py:148); numpy_orig_mask=True returns
    the numpy array of the reference (strict drop-in)."""

    def __init__(self, *args, clamp=-1, process_group=None, numpy_orig_mask=False):
        super().__init__(*args, process_group=process_group)
        self.clamp = clamp
        self.numpy_orig_mask = numpy_orig_mask   # True: strict drop-in, orig_mask as a host numpy array (:640)

    def forward(self, depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1):
        K, ray = self._consts(depth0)
        loss, m0, m1, orig = _FlowConsistency.apply(depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, None, None,
                                                    K, ray, self.clamp, False, self.process_group)
        orig = orig[0, 0]
        return loss, m0, m1, (orig.to('cpu').numpy() if self.numpy_orig_mask else orig)

    tforward = forward


class Multi_Frame_Flow_Consistency_Loss(ProjectionBaseLoss):
    """reference model/networks.py:554-607 (adds the < 1 px reprojection mask from the primary depths)."""

    def __init__(self, *args, clamp=-1, process_group=None):
        super().__init__(*args, process_group=process_group)
        self.clamp = clamp   # stored but unused, as in the reference's fwd (:564-601)

    def forward(self, depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, primary_depth0, primary_depth1):
        K, ray = self._consts(depth0)
        loss, _, _, _ = _FlowConsistency.apply(depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1,
                                               primary_depth0.detach(), primary_depth1.detach(), K, ray, -1.0, True,
                                               self.process_group)
        return loss

    tforward = forward


class DispToDepth(torch.nn.Module):
    """reference model/networks.py:311-319 (plain torch: two elementwise ops feeding the fused geometric loss)."""

    def __init__(self, focal_length, baseline):
        super().__init__()
        self.baseline_focal_length = baseline * focal_length

    def forward(self, disp):
        return self.baseline_focal_length / (torch.nn.functional.relu(disp) + 1e-12)

    tforward = forward
