"""Host-side mirror of the hot-path pieces of the reference's model/multi_frame_networks.py.

    warp(x, flow) -> x_prj                                     reference :83-99
    warp_with_fb_mask(flow_ji, flow_ij) -> (flow_ji warped, mask)   reference :202-209 (gather_warped_xyz)
    resize_like / resize_flow_like / resize_flow_masks_like         reference :42-81

Bilinear resampling with zeros padding and align_corners=True, bit-compatible with the reference's
normalise -> grid_sample round trip, executed by libdis_b200.so.  Gradients flow to x (and to flow when
it requires grad, which the reference never needs).
"""
import torch

from . import _ops


class _FlowWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow):
        out, _, _, _ = _ops.flow_warp_forward(x, flow)
        ctx.save_for_backward(x if ctx.needs_input_grad[1] else None, flow)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, flow = ctx.saved_tensors
        gx, gf = _ops.flow_warp_backward(x, flow, grad_out.contiguous(), want_x=ctx.needs_input_grad[0],
                                         want_flow=ctx.needs_input_grad[1])
        return gx, gf


def warp(x, flow):
    """x [bs,C,h,w], flow [bs,2,h,w] (pixels) -> x sampled at (u + flow_x, v + flow_y)."""
    return _FlowWarp.apply(x, flow)


def warp_with_fb_mask(flow_back, flow_fwd):
    """flow10 = warp(flow_back, flow_fwd) and the forward-backward mask
    |f + f10|^2 < 0.5 + 0.01 (|f|^2 + |f10|^2) as float [bs,1,h,w] -- one fused, gradient-free call
    (the reference computes it under torch.no_grad(), :202-209)."""
    with torch.no_grad():
        out, mask, _, _ = _ops.flow_warp_forward(flow_back, flow_fwd, want_fb_mask=True)
    return out, mask


class _GatherWarped(torch.autograd.Function):
    """[x[tidx], warp(x[j], flow_{tidx,j}) for j != tidx] written straight into the stacked output by one launch
    (no per-warp tensors, no torch.stack copy); backward is one launch into a [tl,bs,C,h,w] gradient."""

    @staticmethod
    def forward(ctx, x, tidx, *flows):
        ctx.tidx = tidx
        ctx.save_for_backward(*flows)
        return _ops.flow_warp_gather_forward(x, flows, tidx)

    @staticmethod
    def backward(ctx, grad_out):
        gx = _ops.flow_warp_gather_backward(ctx.saved_tensors, grad_out.contiguous(), ctx.tidx)
        return (gx, None) + (None,) * len(ctx.saved_tensors)


def gather_warped(x, flow, tidx, with_fb_mask=False):
    """Stack [x[tidx], warp(x[j], flow_{tidx,j}) for j != tidx] -> [tl, bs, C, h, w]: the gather step of
    FuseNet.gather_warped_xyz (reference :187-214) and Block2D3D.gather_warped_feat (:347-360).
    x: [tl, bs, C, h, w] (or a list of tl [bs, C, h, w] tensors); flow: {'flow_ij': [bs, 2, h, w]} at the
    resolution of x.  Gradients flow to x only (the reference detaches / never differentiates the flows).
    with_fb_mask additionally returns the float masks [tl, bs, 1, h, w] (ones for the own frame, the
    forward-backward consistency mask of :202-209 for the others)."""
    if isinstance(x, (list, tuple)):
        x = torch.stack(list(x), dim=0)
    tl = x.shape[0]
    others = [j for j in range(tl) if j != tidx]
    out = _GatherWarped.apply(x, tidx, *[flow[f'flow_{tidx}{j}'].detach() for j in others])
    if not with_fb_mask:
        return out
    masks = [torch.ones_like(x[tidx][:, :1])]
    for j in others:
        masks.append(warp_with_fb_mask(flow[f'flow_{j}{tidx}'], flow[f'flow_{tidx}{j}'])[1])
    return out, torch.stack(masks, dim=0)


class _GatherWarpedAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, keys, *flows):
        ctx.keys = keys
        ctx.save_for_backward(*flows)
        return _ops.flow_warp_gather_all_forward(x, dict(zip(keys, flows)))

    @staticmethod
    def backward(ctx, grad_out):
        gx = _ops.flow_warp_gather_all_backward(dict(zip(ctx.keys, ctx.saved_tensors)), grad_out.contiguous())
        return (gx, None) + (None,) * len(ctx.saved_tensors)


def gather_warped_all(x, flow):
    """gather_warped for every target frame at once: [tl, tl, bs, C, h, w] with out[tidx] == gather_warped(x, flow, tidx),
    i.e. the `warped_feat` stack that Block2D3D.fwd_3d_1 / fwd_3d_2 build in their tidx loops (reference :376-404).
    Two launches forward; the backward pass writes every frame's own-slot gradient and reduces the tl-1 scattered
    ones on top in two launches (no zero-fill pass, no per-tidx gradient accumulation in autograd)."""
    if isinstance(x, (list, tuple)):
        x = torch.stack(list(x), dim=0)
    tl = x.shape[0]
    keys = tuple((i, j) for i in range(tl) for j in range(tl) if i != j)
    return _GatherWarpedAll.apply(x, keys, *[flow[f'flow_{i}{j}'].detach() for (i, j) in keys])


def _target_size(target):
    return (int(target[0]), int(target[1])) if isinstance(target, (tuple, list)) else (target.shape[-2], target.shape[-1])


class _ResizeLike(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x4, size):
        ctx.in_shape = tuple(x4.shape)
        return _ops.resize_bilinear([x4], size, 0)[0]

    @staticmethod
    def backward(ctx, g):
        return _ops.resize_bilinear_backward(g.contiguous(), ctx.in_shape), None


def resize_like(x, target):
    """reference model/multi_frame_networks.py:42-52: bilinear (align_corners=True) resize of the last two dims of x
    ([..., C, H, W]) to those of `target` (a tensor, or an (h, w) pair)."""
    size = _target_size(target)
    lead = x.shape[:-3]
    out = _ResizeLike.apply(x.contiguous().view(-1, *x.shape[-3:]), size)
    return out.view(*lead, *out.shape[1:])


def resize_flow_like(flow, target):
    """reference :54-68: every flow of the dict resized to the target size with its x / y channel rescaled by the width /
    height ratio -- all entries in one launch (the reference: interpolate + 2 in-place multiplies per entry).  A single
    tensor is accepted too.  Flows are data: no gradient."""
    size = _target_size(target)
    with torch.no_grad():
        if isinstance(flow, dict):
            keys = list(flow.keys())
            return dict(zip(keys, _ops.resize_bilinear([flow[k] for k in keys], size, 1)))
        return _ops.resize_bilinear([flow], size, 1)[0]


def resize_flow_masks_like(flow_masks, target):
    """reference :70-81: (interpolate(mask) > 0.5).float() for every mask of the dict, one launch, threshold fused."""
    size = _target_size(target)
    with torch.no_grad():
        if isinstance(flow_masks, dict):
            keys = list(flow_masks.keys())
            return dict(zip(keys, _ops.resize_bilinear([flow_masks[k] for k in keys], size, 2)))
        lead = flow_masks.shape[:-3]
        out = _ops.resize_bilinear([flow_masks.contiguous().view(-1, *flow_masks.shape[-3:])], size, 2)[0]
        return out.view(*lead, *out.shape[1:])


class Conv3DRank:
    """The neighbour selection of one (xyz, mask) pair: Conv3D ranks its candidates by xyz and mask alone (reference
    :469-491), so the layers of a FuseNet level and the checkpoint recompute can share it (conv3d_rank / rank=...)."""

    def __init__(self, xyz_nb, idx, cfg):
        self.xyz_nb, self.idx, self.cfg = xyz_nb, idx, cfg


def conv3d_rank(xyz, mask, ksize=3, stride=1, neighbors=9):
    """-> Conv3DRank for conv3d_gather(..., rank=...).  No gradient is recorded here; conv3d_gather(rank=...) still
    differentiates xyz_neighbors w.r.t. xyz (the selection itself has no gradient)."""
    with torch.no_grad():
        xyz_nb, idx, _ = _ops.conv3d_rank(xyz, mask, ksize, stride, neighbors)
    return Conv3DRank(xyz_nb, idx, (tuple(xyz.shape), ksize, stride, neighbors))


class _Conv3DGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, feat, mask, ksize, stride, neighbors, rank):
        ctx.set_materialize_grads(False)
        if rank is None:
            xyz_nb, feat_nb, idx, _ = _ops.conv3d_gather_forward(xyz, feat, mask, ksize, stride, neighbors)
        else:
            if rank.cfg != (tuple(xyz.shape), ksize, stride, neighbors):
                raise ValueError("rank was computed for a different xyz shape / window / neighbour count")
            idx, xyz_nb = rank.idx.detach(), rank.xyz_nb.detach()   # fresh tensor objects on the cached storage
            feat_nb = _ops.conv3d_gather_features(feat, idx, ksize, stride, neighbors)
        ctx.save_for_backward(idx)
        ctx.cfg = (tuple(feat.shape), ksize, stride, neighbors)
        ctx.mark_non_differentiable(idx)
        return xyz_nb, feat_nb, idx

    @staticmethod
    def backward(ctx, g_xyz_nb, g_feat_nb, _):
        (idx,) = ctx.saved_tensors
        shape, ksize, stride, neighbors = ctx.cfg
        want_xyz = ctx.needs_input_grad[0] and g_xyz_nb is not None
        want_feat = ctx.needs_input_grad[1] and g_feat_nb is not None
        if not (want_xyz or want_feat):
            return None, None, None, None, None, None, None
        g_xyz, g_feat = _ops.conv3d_gather_backward(g_xyz_nb, g_feat_nb, idx, shape, ksize, stride, neighbors,
                                                    want_xyz, want_feat)
        return g_xyz, g_feat, None, None, None, None, None


def conv3d_gather(xyz, feat, mask, ksize=3, stride=1, neighbors=9, rank=None):
    """Neighbour selection + gather of Conv3D.tforward (reference :469-501), everything up to the MLP:
    xyz [tl,bs,3,h,w], feat [tl,bs,C,h,w], mask [tl,bs,1,h,w] ->
      xyz_neighbors [M,neighbors,3] (local coordinates), feat_neighbors [M,neighbors,C],
      neighbors_ind [M,neighbors] (uint8 candidate index (ky*k+kx)*tl + t), M = bs*oh*ow.
    Neighbours come in ascending distance with ties broken by the lowest index (torch.topk(sorted=False) leaves both
    unspecified; the reference only sums over them).  Gradients flow to xyz and feat.
    rank: a Conv3DRank of the same (xyz, mask) from conv3d_rank(): the selection is reused, only the features are gathered."""
    return _Conv3DGather.apply(xyz, feat, mask, ksize, stride, neighbors, rank)
