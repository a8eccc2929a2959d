"""Tensor-level wrappers over the C-ABI: validate, allocate outputs with torch, pass raw device
pointers and the current CUDA stream.  torch is plumbing here (memory + streams); every FLOP of the
path runs in libdis_b200.so.  CPU tensors are rejected: there is no CPU fallback.
"""
import torch

from . import _lib

LOSS_TYPES = {"mse": 0, "sad": 1, "census_mse": 2, "census_sad": 3}


def loss_type_id(type):
    """String -> int map of model/ext_functions.py:142-154 (raises like the reference)."""
    if isinstance(type, int):
        if type not in (0, 1, 2, 3):
            raise Exception("invalid loss type")
        return type
    t = LOSS_TYPES.get(str(type).lower())
    if t is None:
        raise Exception("invalid loss type")
    return t


def _chk(t, name, ndim=4):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: depthinspace_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")
    return t.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _on:
    """Device guard (kernels launch on the tensor's device)."""

    def __init__(self, t):
        self.g = torch.cuda.device(t.device)

    def __enter__(self):
        self.g.__enter__()
        return _lib.load()

    def __exit__(self, *a):
        return self.g.__exit__(*a)


def lcn_forward(x, radius, eps):
    x = _chk(x, "data")
    N, C, H, W = x.shape
    lcn, std = torch.empty_like(x), torch.empty_like(x)
    with _on(x) as lib:
        _lib.check(lib.dis_lcn_forward(_ptr(x), _ptr(lcn), _ptr(std), N * C, H, W, int(radius), float(eps), _stream(x)))
    return lcn, std


def lcn_prepare_input(x, radius, eps):
    """x [bs,tl,1,H,W] -> (im_cat [tl,bs,2,H,W] = cat(LCN(x), x), std [tl,bs,1,H,W]); Worker.copy_data fused."""
    x = _chk(x, "im", 5)
    bs, tl, C, H, W = x.shape
    if C != 1:
        raise ValueError("expected [bs, tl, 1, H, W]")
    im_cat = torch.empty((tl, bs, 2, H, W), dtype=x.dtype, device=x.device)
    std = torch.empty((tl, bs, 1, H, W), dtype=x.dtype, device=x.device)
    with _on(x) as lib:
        _lib.check(lib.dis_lcn_prepare_input(_ptr(x), _ptr(im_cat), _ptr(std), bs, tl, H, W, int(radius), float(eps), _stream(x)))
    return im_cat, std


def photometric_loss_forward(es, ta, block_size, type, eps):
    es, ta = _chk(es, "es"), _chk(ta, "ta")
    if es.shape != ta.shape:
        raise ValueError(f"es {tuple(es.shape)} and ta {tuple(ta.shape)} differ")
    N, C, H, W = es.shape
    out = torch.empty((N, 1, H, W), dtype=es.dtype, device=es.device)
    with _on(es) as lib:
        _lib.check(lib.dis_photometric_loss_forward(_ptr(es), _ptr(ta), _ptr(out), N, C, H, W, int(block_size),
                                                    loss_type_id(type), float(eps), _stream(es)))
    return out


def photometric_loss_backward(es, ta, grad_out, block_size, type, eps):
    es, ta, grad_out = _chk(es, "es"), _chk(ta, "ta"), _chk(grad_out, "grad_out")
    N, C, H, W = es.shape
    if tuple(grad_out.shape) != (N, 1, H, W):
        raise ValueError(f"grad_out must be {(N, 1, H, W)}, got {tuple(grad_out.shape)}")
    grad_es = torch.empty_like(es)
    with _on(es) as lib:
        _lib.check(lib.dis_photometric_loss_backward(_ptr(es), _ptr(ta), _ptr(grad_out), _ptr(grad_es), N, C, H, W,
                                                     int(block_size), loss_type_id(type), float(eps), _stream(es)))
    return grad_es


def pattern_warp(disp, pattern, want_dproj=False, want_corners=False):
    disp = _chk(disp, "disp0")
    N, C, H, W = disp.shape
    pattern = _chk(pattern, "pattern", None).reshape(H, W)
    proj = torch.empty_like(disp)
    dproj = torch.empty_like(disp) if want_dproj else None
    cx = torch.empty(disp.shape, dtype=torch.int32, device=disp.device) if want_corners else None
    cy = torch.empty_like(cx) if want_corners else None
    with _on(disp) as lib:
        _lib.check(lib.dis_pattern_warp_forward(_ptr(disp), _ptr(pattern), _ptr(proj), _ptr(dproj), _ptr(cx), _ptr(cy),
                                                N * C, H, W, _stream(disp)))
    return proj, dproj, cx, cy


def pattern_loss_forward(disp, im, std, pattern, block_size, type, eps, want_proj, want_diff, want_grad, grad_scale=None):
    """-> (out3 [num, den, num/den], proj|None, diff|None, grad_num|None); grad_scale: optional one-element device
    tensor the stored gradient is multiplied by inside the kernel."""
    disp, im = _chk(disp, "disp0"), _chk(im, "im")
    N, C, H, W = disp.shape
    if C != 1 or tuple(im.shape) != (N, 1, H, W):
        raise ValueError(f"disp0 and im must both be [N,1,H,W]; got {tuple(disp.shape)} and {tuple(im.shape)}")
    if std is not None:
        std = _chk(std, "std")
        if std.shape != disp.shape:
            raise ValueError("std must match disp0")
    pattern = _chk(pattern, "pattern", None)
    if pattern.numel() != H * W:
        raise ValueError(f"pattern has {pattern.numel()} elements, expected {H}x{W}")
    new = lambda: torch.empty_like(disp)
    proj = new() if want_proj else None
    diff = new() if want_diff else None
    gnum = new() if want_grad else None
    out3 = torch.empty(3, dtype=torch.float32, device=disp.device)
    with _on(disp) as lib:
        npart = lib.dis_pattern_loss_num_partials(N, H, W)
        partials = torch.empty(2 * max(npart, 1), dtype=torch.float32, device=disp.device)
        s = _stream(disp)
        if grad_scale is not None:
            grad_scale = _chk(grad_scale, "grad_scale", 1)
        _lib.check(lib.dis_pattern_loss_forward_scaled(_ptr(disp), _ptr(im), _ptr(std), _ptr(pattern), _ptr(proj), _ptr(diff),
                                                       _ptr(gnum), _ptr(grad_scale), _ptr(partials), N, H, W,
                                                       int(block_size), loss_type_id(type), float(eps), s))
        _lib.check(lib.dis_reduce_pairs(_ptr(partials), npart, _ptr(out3), s))
    return out3, proj, diff, gnum


def pattern_loss_point_forward(disps, im, std, pattern, block_size, type, want_proj, want_grad, grad_scale=None,
                               workspace=None, reuse_wbox=False):
    """mse / sad pattern loss of S = 1..4 disparity maps of the same frames, no per-pixel map (dis_pattern_loss_point_forward).
    -> (out3 [S,3] rows (num_s, den, num_s/den), list of proj | None, list of grad_num | None)
    workspace: optional float tensor of 2*N*H*W elements; with reuse_wbox its second half already holds the box-filtered
    weights of these frames (same std, block_size)."""
    import ctypes
    S = len(disps)
    disps = [_chk(d, f"disp[{i}]") for i, d in enumerate(disps)]
    im = _chk(im, "im")
    N, C, H, W = disps[0].shape
    if C != 1 or any(d.shape != disps[0].shape for d in disps) or tuple(im.shape) != (N, 1, H, W):
        raise ValueError("all disparity maps and im must be [N,1,H,W] with equal shapes")
    if std is not None:
        std = _chk(std, "std")
        if std.shape != im.shape:
            raise ValueError("std must match im")
    pattern = _chk(pattern, "pattern", None)
    if pattern.numel() != H * W:
        raise ValueError(f"pattern has {pattern.numel()} elements, expected {H}x{W}")
    if workspace is None:
        if reuse_wbox:
            raise ValueError("reuse_wbox needs the workspace of the earlier call")
        workspace = torch.empty(2 * im.numel(), dtype=torch.float32, device=im.device)
    elif workspace.numel() < 2 * im.numel() or workspace.dtype != torch.float32 or not workspace.is_contiguous() or not workspace.is_cuda:
        raise ValueError("workspace must be a contiguous float32 CUDA tensor of 2*N*H*W elements")
    projs = [torch.empty_like(d) for d in disps] if want_proj else None
    grads = [torch.empty_like(d) for d in disps] if want_grad else None
    out3 = torch.empty((S, 3), dtype=torch.float32, device=im.device)
    PtrArr = ctypes.c_void_p * S
    d_arr = PtrArr(*[d.data_ptr() for d in disps])
    p_arr = PtrArr(*[t.data_ptr() for t in projs]) if want_proj else None
    g_arr = PtrArr(*[g.data_ptr() for g in grads]) if want_grad else None
    with _on(im) as lib:
        npart = lib.dis_pattern_loss_point_num_partials(N, H, W)
        partials = torch.empty(2 * S * max(npart, 1), dtype=torch.float32, device=im.device)
        s = _stream(im)
        if grad_scale is not None:
            grad_scale = _chk(grad_scale, "grad_scale", 1)
            if grad_scale.numel() != S:
                raise ValueError(f"grad_scale must have {S} elements")
        _lib.check(lib.dis_pattern_loss_point_forward(d_arr, S, _ptr(im), _ptr(std), _ptr(pattern), p_arr, g_arr, _ptr(grad_scale),
                                                      _ptr(workspace), int(bool(reuse_wbox)), _ptr(partials), N, H, W,
                                                      int(block_size), loss_type_id(type), s))
        _lib.check(lib.dis_reduce_pairs_batched(_ptr(partials), npart, S, _ptr(out3), s))
    return out3, projs, grads


def pattern_loss_multi_forward(disps, im, std, pattern, block_size, type, eps, want_grad, grad_scale=None):
    """S = 2 or 4 disparity maps of the same frames (census types; mse / sad: S = 1..4, through the point-wise path).
    -> (out3 [S,3] rows (num_s, den, num_s/den), list of grad_num | None)
    grad_scale: optional device tensor of S floats; grad_num[s] is multiplied by it inside the kernel (final gradients)."""
    import ctypes
    S = len(disps)
    if loss_type_id(type) < 2:
        out3, _, grads = pattern_loss_point_forward(disps, im, std, pattern, block_size, type, False, want_grad, grad_scale)
        return out3, grads
    disps = [_chk(d, f"disp[{i}]") for i, d in enumerate(disps)]
    im = _chk(im, "im")
    N, C, H, W = disps[0].shape
    if C != 1 or any(d.shape != disps[0].shape for d in disps) or tuple(im.shape) != (N, 1, H, W):
        raise ValueError("all disparity maps and im must be [N,1,H,W] with equal shapes")
    if std is not None:
        std = _chk(std, "std")
        if std.shape != im.shape:
            raise ValueError("std must match im")
    pattern = _chk(pattern, "pattern", None)
    if pattern.numel() != H * W:
        raise ValueError(f"pattern has {pattern.numel()} elements, expected {H}x{W}")
    grads = [torch.empty_like(d) for d in disps] if want_grad else None
    out3 = torch.empty((S, 3), dtype=torch.float32, device=im.device)
    PtrArr = ctypes.c_void_p * S
    d_arr = PtrArr(*[d.data_ptr() for d in disps])
    g_arr = PtrArr(*[g.data_ptr() for g in grads]) if want_grad else None
    with _on(im) as lib:
        npart = lib.dis_pattern_loss_multi_num_partials(N, H, W)
        partials = torch.empty(2 * S * max(npart, 1), dtype=torch.float32, device=im.device)
        s = _stream(im)
        if grad_scale is not None:
            grad_scale = _chk(grad_scale, "grad_scale", 1)
            if grad_scale.numel() != S:
                raise ValueError(f"grad_scale must have {S} elements")
        _lib.check(lib.dis_pattern_loss_multi_forward_scaled(d_arr, S, _ptr(im), _ptr(std), _ptr(pattern), g_arr,
                                                             _ptr(grad_scale), _ptr(partials), N, H, W, int(block_size),
                                                             loss_type_id(type), float(eps), s))
        _lib.check(lib.dis_reduce_pairs_batched(_ptr(partials), npart, S, _ptr(out3), s))
    return out3, grads


def scale_by_device_scalar(x, numer, denom=None):
    """x * numer / denom with numer, denom one-element device tensors (no host sync)."""
    x = x.contiguous()
    out = torch.empty_like(x)
    numer = numer.reshape(1).to(torch.float32).contiguous()
    with _on(x) as lib:
        _lib.check(lib.dis_scale_by_device_scalar(_ptr(x), _ptr(out), x.numel(), _ptr(numer), _ptr(denom), _stream(x)))
    return out


def abs_sum(a):
    """-> out3 [sum|a|, count, mean] (device tensor, no host sync); the sigma normaliser of the photometric terms."""
    return l1_forward(a, None, False)[0]


def l1_forward(a, b, want_grad):
    """-> (out3 [sum|a-b|, count, mean], sign(a-b) | None); b = None means 0"""
    a = _chk(a, "a", None)
    b = _chk(b, "b", None) if b is not None else None
    if b is not None and a.shape != b.shape:
        b = b.expand_as(a).contiguous()
    sgn = torch.empty_like(a) if want_grad else None
    out3 = torch.empty(3, dtype=torch.float32, device=a.device)
    with _on(a) as lib:
        npart = lib.dis_l1_num_partials(a.numel())
        partials = torch.empty(2 * npart, dtype=torch.float32, device=a.device)
        s = _stream(a)
        _lib.check(lib.dis_l1_forward(_ptr(a), _ptr(b), _ptr(sgn), _ptr(partials), a.numel(), s))
        _lib.check(lib.dis_reduce_pairs(_ptr(partials), npart, _ptr(out3), s))
    return out3, sgn


def masked_l1_forward(a, b, noise, threshold, want_grad):
    """-> (out3 [sum |a - b + noise| * valid, sum valid, ratio], sign * valid | None), valid = (b > threshold)"""
    a, b = _chk(a, "a", None), _chk(b, "b", None)
    if a.shape != b.shape:
        b = b.expand_as(a).contiguous()
    if noise is not None:
        noise = _chk(noise, "noise", None)
        if noise.shape != a.shape:
            raise ValueError("noise must have the shape of a")
    sgn = torch.empty_like(a) if want_grad else None
    out3 = torch.empty(3, dtype=torch.float32, device=a.device)
    with _on(a) as lib:
        npart = lib.dis_l1_num_partials(a.numel())
        partials = torch.empty(2 * npart, dtype=torch.float32, device=a.device)
        s = _stream(a)
        _lib.check(lib.dis_masked_l1_forward(_ptr(a), _ptr(b), _ptr(noise), float(threshold), _ptr(sgn), _ptr(partials),
                                             a.numel(), s))
        _lib.check(lib.dis_reduce_pairs(_ptr(partials), npart, _ptr(out3), s))
    return out3, sgn


def mul(a, b):
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty_like(a)
    with _on(a) as lib:
        _lib.check(lib.dis_mul(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream(a)))
    return out


def sobel_forward(x, ksize):
    x = _chk(x, "x")
    N, C, H, W = x.shape
    if C != 1:
        raise ValueError("SobelFilter expects [N,1,H,W]")
    out = torch.empty((N, 2, H, W), dtype=x.dtype, device=x.device)
    with _on(x) as lib:
        _lib.check(lib.dis_sobel_forward(_ptr(x), _ptr(out), N, H, W, int(ksize), _stream(x)))
    return out


def sobel_backward(grad_out, ksize):
    grad_out = _chk(grad_out, "grad_out")
    N, _, H, W = grad_out.shape
    gx = torch.empty((N, 1, H, W), dtype=grad_out.dtype, device=grad_out.device)
    with _on(grad_out) as lib:
        _lib.check(lib.dis_sobel_backward(_ptr(grad_out), _ptr(gx), N, H, W, int(ksize), _stream(grad_out)))
    return gx


def smooth_loss_buffers(disp, want_grad):
    """Outputs of smooth_loss_forward allocated ahead of the launch -> (out3, gsum | None, partials)."""
    disp = _chk(disp, "disp")
    N, _, H, W = disp.shape
    with _on(disp) as lib:
        npart = lib.dis_smooth_loss_num_partials(N, H, W)
    return (torch.empty(3, dtype=torch.float32, device=disp.device), torch.empty_like(disp) if want_grad else None,
            torch.empty(2 * max(npart, 1), dtype=torch.float32, device=disp.device))


def smooth_loss_forward(disp, im, want_grad, grad_scale=1.0, accumulate_into=None, launch_stream=None, buffers=None):
    """-> (out3 [sum, count, mean], grad_sum * grad_scale | None); accumulate_into: an existing gradient tensor of disp's
    shape that the scaled gradient is ADDED to (returned in place of a fresh tensor).
    launch_stream + buffers: launch on another torch.cuda.Stream than the current one, into outputs from
    smooth_loss_buffers().  The buffers must have been allocated BEFORE the point of the current stream that
    launch_stream waits for: memory handed out later may still be in use by work queued on the current stream."""
    disp, im = _chk(disp, "disp"), _chk(im, "im")
    if disp.shape != im.shape or disp.shape[1] != 1:
        raise ValueError(f"disp and im must both be [N,1,H,W]; got {tuple(disp.shape)} and {tuple(im.shape)}")
    if (launch_stream is None) != (buffers is None):
        raise ValueError("launch_stream and buffers go together")
    N, _, H, W = disp.shape
    if accumulate_into is not None:
        gsum = _chk(accumulate_into, "accumulate_into")
        if gsum.shape != disp.shape or gsum.data_ptr() != accumulate_into.data_ptr():
            raise ValueError("accumulate_into must be a contiguous tensor of disp's shape")
    elif buffers is not None:
        gsum = buffers[1]
        if want_grad and gsum is None:
            raise ValueError("buffers were allocated without a gradient plane")
        if not want_grad:
            gsum = None
    else:
        gsum = torch.empty_like(disp) if want_grad else None
    out3 = buffers[0] if buffers is not None else torch.empty(3, dtype=torch.float32, device=disp.device)
    with _on(disp) as lib:
        npart = lib.dis_smooth_loss_num_partials(N, H, W)
        partials = buffers[2] if buffers is not None else torch.empty(2 * max(npart, 1), dtype=torch.float32, device=disp.device)
        if partials.numel() < 2 * max(npart, 1):
            raise ValueError("partials buffer too small")
        s = _stream(disp) if launch_stream is None else launch_stream.cuda_stream
        _lib.check(lib.dis_smooth_loss_forward_scaled(_ptr(disp), _ptr(im), _ptr(gsum), _ptr(partials), N, H, W,
                                                      float(grad_scale), int(accumulate_into is not None), s))
        _lib.check(lib.dis_reduce_pairs(_ptr(partials), npart, _ptr(out3), s))
    return out3, gsum


def flow_warp_forward(x, flow, want_fb_mask=False, want_corners=False):
    x, flow = _chk(x, "x"), _chk(flow, "flow")
    N, C, H, W = x.shape
    if tuple(flow.shape) != (N, 2, H, W):
        raise ValueError(f"flow must be {(N, 2, H, W)}, got {tuple(flow.shape)}")
    out = torch.empty_like(x)
    mask = torch.empty((N, 1, H, W), dtype=x.dtype, device=x.device) if want_fb_mask else None
    cx = torch.empty((N, 1, H, W), dtype=torch.int32, device=x.device) if want_corners else None
    cy = torch.empty_like(cx) if want_corners else None
    with _on(x) as lib:
        _lib.check(lib.dis_flow_warp_forward(_ptr(x), _ptr(flow), _ptr(out), _ptr(mask), _ptr(cx), _ptr(cy), N, C, H, W,
                                             _stream(x)))
    return out, mask, cx, cy


def flow_warp_backward(x, flow, grad_out, want_x=True, want_flow=False):
    flow, grad_out = _chk(flow, "flow"), _chk(grad_out, "grad_out")
    N, C, H, W = grad_out.shape
    x = _chk(x, "x") if x is not None else None
    gx = torch.empty_like(grad_out) if want_x else None
    gf = torch.empty_like(flow) if want_flow else None
    with _on(grad_out) as lib:
        _lib.check(lib.dis_flow_warp_backward(_ptr(x), _ptr(flow), _ptr(grad_out), _ptr(gx), _ptr(gf), N, C, H, W,
                                              _stream(grad_out)))
    return gx, gf


def _flow_ptr_array(flows, tl, bs, H, W):
    import ctypes
    if len(flows) != tl - 1:
        raise ValueError(f"expected {tl - 1} flows, got {len(flows)}")
    flows = [_chk(f, f"flow[{i}]") for i, f in enumerate(flows)]
    for f in flows:
        if tuple(f.shape) != (bs, 2, H, W):
            raise ValueError(f"every flow must be {(bs, 2, H, W)}, got {tuple(f.shape)}")
    arr = (ctypes.c_void_p * max(tl - 1, 1))(*[f.data_ptr() for f in flows])
    return flows, arr


def flow_warp_gather_forward(x, flows, tidx):
    """x [tl,bs,C,h,w]; flows: the tl-1 tensors flow_{tidx,j}, j != tidx in increasing j -> out [tl,bs,C,h,w]."""
    x = _chk(x, "x", 5)
    tl, bs, C, H, W = x.shape
    flows, arr = _flow_ptr_array(flows, tl, bs, H, W)
    out = torch.empty_like(x)
    with _on(x) as lib:
        _lib.check(lib.dis_flow_warp_gather_forward(_ptr(x), arr, _ptr(out), tl, int(tidx), bs, C, H, W, _stream(x)))
    return out


def flow_warp_gather_backward(flows, grad_out, tidx):
    grad_out = _chk(grad_out, "grad_out", 5)
    tl, bs, C, H, W = grad_out.shape
    flows, arr = _flow_ptr_array(flows, tl, bs, H, W)
    gx = torch.empty_like(grad_out)
    with _on(grad_out) as lib:
        _lib.check(lib.dis_flow_warp_gather_backward(arr, _ptr(grad_out), _ptr(gx), tl, int(tidx), bs, C, H, W,
                                                     _stream(grad_out)))
    return gx


def _flow_matrix_ptr_array(flows, tl, bs, H, W):
    """flows: dict {(i, j): tensor} for all i != j -> ctypes array of tl*tl pointers (diagonal NULL) + keep-alive list"""
    import ctypes
    keep, ptrs = [], []
    for i in range(tl):
        for j in range(tl):
            if i == j:
                ptrs.append(None)
                continue
            f = _chk(flows[(i, j)], f"flow_{i}{j}")
            if tuple(f.shape) != (bs, 2, H, W):
                raise ValueError(f"flow_{i}{j} must be {(bs, 2, H, W)}, got {tuple(f.shape)}")
            keep.append(f)
            ptrs.append(f.data_ptr())
    return keep, (ctypes.c_void_p * (tl * tl))(*ptrs)


def flow_warp_gather_all_forward(x, flows):
    """x [tl,bs,C,h,w]; flows {(i, j): flow_ij [bs,2,h,w]} -> out [tl,tl,bs,C,h,w], out[i] = gather for target frame i."""
    x = _chk(x, "x", 5)
    tl, bs, C, H, W = x.shape
    keep, arr = _flow_matrix_ptr_array(flows, tl, bs, H, W)
    out = torch.empty((tl, tl, bs, C, H, W), dtype=x.dtype, device=x.device)
    with _on(x) as lib:
        _lib.check(lib.dis_flow_warp_gather_all_forward(_ptr(x), arr, _ptr(out), tl, bs, C, H, W, _stream(x)),
                   launches=2 if tl > 1 else 1)
    return out


def flow_warp_gather_all_backward(flows, grad_out):
    grad_out = _chk(grad_out, "grad_out", 6)
    tl, tl2, bs, C, H, W = grad_out.shape
    if tl != tl2:
        raise ValueError("grad_out must be [tl,tl,bs,C,h,w]")
    keep, arr = _flow_matrix_ptr_array(flows, tl, bs, H, W)
    gx = torch.empty((tl, bs, C, H, W), dtype=grad_out.dtype, device=grad_out.device)
    with _on(grad_out) as lib:
        _lib.check(lib.dis_flow_warp_gather_all_backward(arr, _ptr(grad_out), _ptr(gx), tl, bs, C, H, W, _stream(grad_out)),
                   launches=2 if tl > 1 else 1)
    return gx


def _chk_long(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: depthinspace_b200 has no CPU path")
    if t.dtype != torch.int64:
        raise TypeError(f"{name} must be int64, got {t.dtype}")
    return t.contiguous()


def ext_nn(in0, in1):
    """ext_cuda.nn_cuda (reference model/ext_functions.py:46): in0 [n0,D], in1 [n1,D] -> int64 [n0] nearest row of in1."""
    in0, in1 = _chk(in0, "in0", 2), _chk(in1, "in1", 2)
    if in0.shape[1] != in1.shape[1]:
        raise ValueError("in0 and in1 must have the same number of columns")
    out = torch.empty(in0.shape[0], dtype=torch.int64, device=in0.device)
    with _on(in0) as lib:
        _lib.check(lib.dis_ext_nn(_ptr(in0), _ptr(in1), out.data_ptr(), in0.shape[0], in1.shape[0], in0.shape[1], _stream(in0)))
    return out


def ext_crosscheck(in0, in1):
    """ext_cuda.crosscheck_cuda (:64): in0 int64 [n0] -> in1, in1 int64 [n1] -> in0; uint8 [n0] = (in1[in0[i]] == i)."""
    in0, in1 = _chk_long(in0, "in0"), _chk_long(in1, "in1")
    out = torch.empty(in0.numel(), dtype=torch.uint8, device=in0.device)
    with _on(in0) as lib:
        _lib.check(lib.dis_ext_crosscheck(in0.data_ptr(), in1.data_ptr(), out.data_ptr(), in0.numel(), in1.numel(), _stream(in0)))
    return out.view(in0.shape)


def ext_proj_nn(xyz0, xyz1, K, patch_size):
    """ext_cuda.proj_nn_cuda (:81): xyz0, xyz1 [bs,H,W,3], K [3,3] -> int64 [bs,H,W] (flat index into xyz1 or -1)."""
    xyz0, xyz1, K = _chk(xyz0, "xyz0"), _chk(xyz1, "xyz1"), _chk(K, "K", 2)
    bs, H, W, three = xyz0.shape
    if three != 3 or xyz1.shape != xyz0.shape or tuple(K.shape) != (3, 3):
        raise ValueError("xyz0, xyz1 must be [bs,H,W,3] with equal shapes and K [3,3]")
    out = torch.empty((bs, H, W), dtype=torch.int64, device=xyz0.device)
    with _on(xyz0) as lib:
        _lib.check(lib.dis_ext_proj_nn(_ptr(xyz0), _ptr(xyz1), _ptr(K), out.data_ptr(), bs, H, W, int(patch_size), _stream(xyz0)))
    return out


def ext_xcorrvol(in0, in1, n_disps, block_size):
    """ext_cuda.xcorrvol_cuda (:100): in0, in1 [C,H,W] -> [n_disps,H,W] zero-normalised block cross correlation."""
    in0, in1 = _chk(in0, "in0", 3), _chk(in1, "in1", 3)
    if in0.shape != in1.shape:
        raise ValueError("in0 and in1 must have equal shapes")
    C, H, W = in0.shape
    out = torch.empty((int(n_disps), H, W), dtype=torch.float32, device=in0.device)
    with _on(in0) as lib:
        _lib.check(lib.dis_ext_xcorrvol(_ptr(in0), _ptr(in1), _ptr(out), C, H, W, int(n_disps), int(block_size), _stream(in0)))
    return out


def resize_bilinear(tensors, size, mode=0):
    """Bilinear align_corners=True resize of a list of equally shaped [N,C,H,W] tensors to `size` in one launch.
    mode 0 plain, 1 flow (x / y channel rescaled by the size ratio), 2 mask (> 0.5 -> 1 / 0).  -> list of outputs"""
    import ctypes
    tensors = [_chk(t, f"tensor[{i}]") for i, t in enumerate(tensors)]
    if not tensors:
        return []
    N, C, H, W = tensors[0].shape
    if any(t.shape != tensors[0].shape for t in tensors):
        raise ValueError("all tensors of one resize call must have the same shape")
    oh, ow = int(size[0]), int(size[1])
    outs = [torch.empty((N, C, oh, ow), dtype=t.dtype, device=t.device) for t in tensors]
    Arr = ctypes.c_void_p * len(tensors)
    with _on(tensors[0]) as lib:
        _lib.check(lib.dis_resize_bilinear_forward(Arr(*[t.data_ptr() for t in tensors]), Arr(*[o.data_ptr() for o in outs]),
                                                   len(tensors), N, C, H, W, oh, ow, int(mode), _stream(tensors[0])),
                   launches=(len(tensors) + 55) // 56)
    return outs


def resize_bilinear_backward(grad_out, in_shape):
    grad_out = _chk(grad_out, "grad_out")
    N, C, oh, ow = grad_out.shape
    H, W = int(in_shape[-2]), int(in_shape[-1])
    g = torch.empty((N, C, H, W), dtype=grad_out.dtype, device=grad_out.device)
    with _on(grad_out) as lib:
        _lib.check(lib.dis_resize_bilinear_backward(_ptr(grad_out), _ptr(g), N, C, H, W, oh, ow, _stream(grad_out)))
    return g


def lcn_backward(data, lcn, std, g_lcn, g_std, radius, eps):
    data, lcn, std = _chk(data, "data"), _chk(lcn, "lcn"), _chk(std, "std")
    N, C, H, W = data.shape
    g_lcn = _chk(g_lcn, "g_lcn") if g_lcn is not None else None
    g_std = _chk(g_std, "g_std") if g_std is not None else None
    if g_lcn is None and g_std is None:
        return torch.zeros_like(data)
    out = torch.empty_like(data)
    work = torch.empty(2 * data.numel(), dtype=torch.float32, device=data.device)
    with _on(data) as lib:
        _lib.check(lib.dis_lcn_backward(_ptr(data), _ptr(lcn), _ptr(std), _ptr(g_lcn), _ptr(g_std), _ptr(out), _ptr(work),
                                        N * C, H, W, int(radius), float(eps), _stream(data)), launches=2)
    return out


def flow_consistency_dir(depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, K, ray, clamp, primary_depth1,
                         want_mask, want_orig_mask, want_grad0, want_grad1, fb_scale=0.02):
    """One direction of the flow-consistency loss.
    -> (out3 [sum(diff*mask), sum(mask), ratio], mask|None, orig_mask|None, grad_depth0|None, grad_depth1|None)"""
    depth0, depth1 = _chk(depth0, "depth0"), _chk(depth1, "depth1")
    flow0, flow1, amb0, amb1 = _chk(flow0, "flow0"), _chk(flow1, "flow1"), _chk(amb0, "amb0"), _chk(amb1, "amb1")
    bs, C, H, W = depth0.shape
    if C != 1 or depth1.shape != depth0.shape or tuple(flow0.shape) != (bs, 2, H, W) or flow1.shape != flow0.shape \
            or amb0.shape != amb1.shape or amb0.shape[0] != bs or tuple(amb0.shape[-2:]) != (H, W):
        raise ValueError("flow-consistency loss: inconsistent shapes")
    R0, R1 = _chk(R0, "R0", 3), _chk(R1, "R1", 3)
    t0, t1 = _chk(t0, "t0", None).reshape(bs, 3), _chk(t1, "t1", None).reshape(bs, 3)
    K, ray = _chk(K, "K", None).reshape(3, 3), _chk(ray, "ray", None).reshape(H * W, 3)
    if primary_depth1 is not None:
        primary_depth1 = _chk(primary_depth1, "primary_depth1")
    new = lambda: torch.empty_like(depth0)
    mask = new() if want_mask else None
    orig = new() if want_orig_mask else None
    g0 = new() if want_grad0 else None
    g1 = new() if want_grad1 else None
    out3 = torch.empty(3, dtype=torch.float32, device=depth0.device)
    with _on(depth0) as lib:
        npart = lib.dis_flow_consistency_num_partials(bs, H, W)
        partials = torch.empty(2 * max(npart, 1), dtype=torch.float32, device=depth0.device)
        s = _stream(depth0)
        _lib.check(lib.dis_flow_consistency_forward(_ptr(depth0), _ptr(depth1), _ptr(R0), _ptr(t0), _ptr(R1), _ptr(t1),
                                                    _ptr(flow0), _ptr(flow1), _ptr(amb0), _ptr(amb1), int(amb0.shape[1]),
                                                    _ptr(primary_depth1), _ptr(K), _ptr(ray), float(clamp), float(fb_scale),
                                                    _ptr(mask), _ptr(orig), _ptr(g0), _ptr(g1), _ptr(partials), bs, H, W, s))
        _lib.check(lib.dis_reduce_pairs(_ptr(partials), npart, _ptr(out3), s))
    return out3, mask, orig, g0, g1


def combine2(a, b, numer, den_a, den_b, eps):
    """numer * (a / (den_a + eps) + b / (den_b + eps)) with one-element device tensors (no host sync)."""
    a = a.contiguous()
    b = b.contiguous() if b is not None else None
    out = torch.empty_like(a)
    numer = numer.reshape(1).to(torch.float32).contiguous()
    with _on(a) as lib:
        _lib.check(lib.dis_combine2(_ptr(a), _ptr(b), _ptr(out), a.numel(), _ptr(numer), _ptr(den_a), _ptr(den_b),
                                    float(eps), _stream(a)))
    return out


def geometric_grad_combine(planes, frame_of, scale, disp, baseline_focal):
    """planes: list of [bs,1,H,W] gradient planes, frame_of: their frame indices, scale: device tensor [len(planes)],
    disp [tl,bs,1,H,W] -> grad_disp [tl,bs,1,H,W] (see dis_geometric_grad_combine)."""
    import ctypes
    disp = _chk(disp, "disp", 5)
    tl, bs, C, H, W = disp.shape
    planes = [_chk(p, f"plane[{i}]") for i, p in enumerate(planes)]
    for p in planes:
        if tuple(p.shape) != (bs, 1, H, W):
            raise ValueError(f"every gradient plane must be {(bs, 1, H, W)}, got {tuple(p.shape)}")
    n = len(planes)
    scale = _chk(scale, "scale", 1)
    if C != 1 or scale.numel() != n or len(frame_of) != n:
        raise ValueError("geometric_grad_combine: inconsistent arguments")
    out = torch.empty_like(disp)
    arr = (ctypes.c_void_p * max(n, 1))(*[p.data_ptr() for p in planes])
    fo = (ctypes.c_int * max(n, 1))(*[int(f) for f in frame_of])
    with _on(disp) as lib:
        _lib.check(lib.dis_geometric_grad_combine(arr, fo, n, _ptr(scale), _ptr(disp), float(baseline_focal), _ptr(out),
                                                  tl, bs, H, W, _stream(disp)))
    return out


def conv3d_gather_forward(xyz, feat, mask, ksize, stride, neighbors):
    """-> (xyz_nb [M,nb,3], feat_nb [M,nb,C], idx uint8 [M,nb], (oh, ow))"""
    xyz, feat, mask = _chk(xyz, "xyz", 5), _chk(feat, "feat", 5), _chk(mask, "mask", 5)
    tl, bs, C, h, w = feat.shape
    if tuple(xyz.shape) != (tl, bs, 3, h, w) or tuple(mask.shape) != (tl, bs, 1, h, w):
        raise ValueError("expected xyz [tl,bs,3,h,w], feat [tl,bs,C,h,w], mask [tl,bs,1,h,w]")
    with _on(xyz) as lib:
        oh, ow = lib.dis_conv3d_out_size(h, ksize, stride), lib.dis_conv3d_out_size(w, ksize, stride)
        M = bs * oh * ow
        xyz_nb = torch.empty((M, neighbors, 3), dtype=torch.float32, device=xyz.device)
        feat_nb = torch.empty((M, neighbors, C), dtype=torch.float32, device=xyz.device)
        idx = torch.empty((M, neighbors), dtype=torch.uint8, device=xyz.device)
        scratch = torch.empty(lib.dis_conv3d_scratch_elems(tl, bs, h, w), dtype=torch.float32, device=xyz.device)
        _lib.check(lib.dis_conv3d_gather_forward(_ptr(xyz), _ptr(feat), _ptr(mask), _ptr(xyz_nb), _ptr(feat_nb), _ptr(idx),
                                                 _ptr(scratch), tl, bs, C, h, w, int(ksize), int(stride), int(neighbors),
                                                 _stream(xyz)), launches=4)
    return xyz_nb, feat_nb, idx, (oh, ow)


def conv3d_rank(xyz, mask, ksize, stride, neighbors):
    """Neighbour selection only -> (xyz_nb [M,nb,3], idx uint8 [M,nb], (oh, ow)); depends on xyz and mask, not on features."""
    xyz, mask = _chk(xyz, "xyz", 5), _chk(mask, "mask", 5)
    tl, bs, _, h, w = xyz.shape
    if xyz.shape[2] != 3 or tuple(mask.shape) != (tl, bs, 1, h, w):
        raise ValueError("expected xyz [tl,bs,3,h,w], mask [tl,bs,1,h,w]")
    with _on(xyz) as lib:
        oh, ow = lib.dis_conv3d_out_size(h, ksize, stride), lib.dis_conv3d_out_size(w, ksize, stride)
        M = bs * oh * ow
        xyz_nb = torch.empty((M, neighbors, 3), dtype=torch.float32, device=xyz.device)
        idx = torch.empty((M, neighbors), dtype=torch.uint8, device=xyz.device)
        scratch = torch.empty(lib.dis_conv3d_scratch_elems(tl, bs, h, w), dtype=torch.float32, device=xyz.device)
        _lib.check(lib.dis_conv3d_rank(_ptr(xyz), _ptr(mask), _ptr(xyz_nb), _ptr(idx), _ptr(scratch), tl, bs, h, w, int(ksize),
                                       int(stride), int(neighbors), _stream(xyz)), launches=3)
    return xyz_nb, idx, (oh, ow)


def conv3d_gather_features(feat, idx, ksize, stride, neighbors):
    """Feature gather for a given selection -> feat_nb [M,nb,C]."""
    feat = _chk(feat, "feat", 5)
    tl, bs, C, h, w = feat.shape
    if idx.dtype != torch.uint8 or not idx.is_cuda or not idx.is_contiguous():
        raise TypeError("idx must be a contiguous uint8 CUDA tensor")
    with _on(feat) as lib:
        oh, ow = lib.dis_conv3d_out_size(h, ksize, stride), lib.dis_conv3d_out_size(w, ksize, stride)
        M = bs * oh * ow
        if tuple(idx.shape) != (M, neighbors):
            raise ValueError(f"idx must be [{M},{neighbors}], got {tuple(idx.shape)}")
        feat_nb = torch.empty((M, neighbors, C), dtype=torch.float32, device=feat.device)
        _lib.check(lib.dis_conv3d_gather_features(_ptr(feat), _ptr(idx), _ptr(feat_nb), tl, bs, C, h, w, int(ksize), int(stride),
                                                  int(neighbors), _stream(feat)))
    return feat_nb


def conv3d_gather_backward(g_xyz_nb, g_feat_nb, idx, shape, ksize, stride, neighbors, want_xyz, want_feat):
    tl, bs, C, h, w = shape
    g_xyz = torch.empty((tl, bs, 3, h, w), dtype=torch.float32, device=idx.device) if want_xyz else None
    g_feat = torch.empty((tl, bs, C, h, w), dtype=torch.float32, device=idx.device) if want_feat else None
    g_xyz_nb = g_xyz_nb.contiguous() if g_xyz_nb is not None else None
    g_feat_nb = g_feat_nb.contiguous() if g_feat_nb is not None else None
    with _on(idx) as lib:
        _lib.check(lib.dis_conv3d_gather_backward(_ptr(g_xyz_nb), _ptr(g_feat_nb), _ptr(idx), _ptr(g_xyz), _ptr(g_feat),
                                                  tl, bs, C, h, w, int(ksize), int(stride), int(neighbors), _stream(idx)))
    return g_xyz, g_feat
