"""ctypes front-end of the plain-C oracle (oracle/dis_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under depthinspace_b200/
may import this module.

All functions take / return numpy arrays; ``prec`` selects the fp32 build (same op
order as the CUDA path) or the fp64 build (tie-breaker).  Frames are independent, so
the batch is fanned out over a thread pool (ctypes drops the GIL) -- the C code itself
is scalar and single-threaded.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}
TYPES = {"mse": 0, "sad": 1, "census_mse": 2, "census_sad": 3}


def build(force=False):
    """Compile both oracle libraries with gcc (idempotent)."""
    outs = [os.path.join(_HERE, "_build", f"libdis_oracle_{p}.so") for p in ("f32", "f64")]
    src = os.path.join(_HERE, "dis_oracle.c")
    stale = force or any(not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(src) for o in outs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
    return outs


def _lib(prec):
    if prec not in _LIBS:
        build()
        lib = ctypes.CDLL(os.path.join(_HERE, "_build", f"libdis_oracle_{prec}.so"))
        assert lib.orc_sizeof_real() == (4 if prec == "f32" else 8)
        lib.orc_smooth_loss.restype = ctypes.c_double
        _LIBS[prec] = lib
    return _LIBS[prec]


def _dt(prec):
    return np.float32 if prec == "f32" else np.float64


def _real(prec, v):
    return ctypes.c_float(v) if prec == "f32" else ctypes.c_double(v)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, prec):
    return None if a is None else np.ascontiguousarray(a, dtype=_dt(prec))


def _threads():
    return max(1, min(32, os.cpu_count() or 1))


def _fan(n, fn):
    """Run fn(i) for i in range(n) on a thread pool; returns list of results."""
    if n <= 1:
        return [fn(i) for i in range(n)]
    with ThreadPoolExecutor(_threads()) as ex:
        return list(ex.map(fn, range(n)))


def lcn_forward(x, radius, eps, prec="f32"):
    x = _c(x, prec)
    N, _, H, W = x.shape
    lcn, std = np.empty_like(x), np.empty_like(x)
    lib = _lib(prec)
    _fan(N, lambda i: lib.orc_lcn_forward(_p(x[i]), _p(lcn[i]), _p(std[i]), 1, H, W, int(radius), _real(prec, eps)))
    return lcn, std


def photometric_forward(es, ta, block_size, type, eps, prec="f32"):
    es, ta = _c(es, prec), _c(ta, prec)
    N, C, H, W = es.shape
    out = np.empty((N, 1, H, W), _dt(prec))
    lib = _lib(prec)
    rcs = _fan(N, lambda i: lib.orc_photometric_forward(_p(es[i]), _p(ta[i]), _p(out[i]), 1, C, H, W,
                                                       int(block_size), int(type), _real(prec, eps)))
    if any(rcs):
        raise Exception("invalid loss type")
    return out


def photometric_backward(es, ta, grad_out, block_size, type, eps, prec="f32"):
    es, ta, grad_out = _c(es, prec), _c(ta, prec), _c(grad_out, prec)
    N, C, H, W = es.shape
    g = np.empty_like(es)
    lib = _lib(prec)
    rcs = _fan(N, lambda i: lib.orc_photometric_backward(_p(es[i]), _p(ta[i]), _p(grad_out[i]), _p(g[i]), 1, C, H, W,
                                                        int(block_size), int(type), _real(prec, eps)))
    if any(rcs):
        raise Exception("invalid loss type")
    return g


def pattern_warp(disp, pattern, prec="f32"):
    """-> (pattern_proj [N,1,H,W], d proj/d disp, x0 int32, y0 int32)."""
    disp, pattern = _c(disp, prec), _c(pattern, prec)
    N, _, H, W = disp.shape
    pat = pattern.reshape(H, W)
    proj, dpd = np.empty_like(disp), np.empty_like(disp)
    ix0, iy0 = np.empty(disp.shape, np.int32), np.empty(disp.shape, np.int32)
    lib = _lib(prec)
    _fan(N, lambda i: lib.orc_pattern_warp(_p(disp[i]), _p(pat), _p(proj[i]), _p(dpd[i]), _p(ix0[i]), _p(iy0[i]), 1, H, W))
    return proj, dpd, ix0, iy0


def flow_warp_forward(x, flow, prec="f32"):
    x, flow = _c(x, prec), _c(flow, prec)
    N, C, H, W = x.shape
    out = np.empty_like(x)
    ix0, iy0 = np.empty((N, 1, H, W), np.int32), np.empty((N, 1, H, W), np.int32)
    lib = _lib(prec)
    _fan(N, lambda i: lib.orc_flow_warp_forward(_p(x[i]), _p(flow[i]), _p(out[i]), _p(ix0[i]), _p(iy0[i]), 1, C, H, W))
    return out, ix0, iy0


def flow_warp_backward(x, flow, grad_out, want_flow_grad=False, prec="f32"):
    x, flow, grad_out = _c(x, prec), _c(flow, prec), _c(grad_out, prec)
    N, C, H, W = x.shape
    gx = np.empty_like(x)
    gf = np.empty_like(flow) if want_flow_grad else None
    lib = _lib(prec)
    _fan(N, lambda i: lib.orc_flow_warp_backward(_p(x[i]), _p(flow[i]), _p(grad_out[i]), _p(gx[i]),
                                                 _p(gf[i]) if want_flow_grad else None, 1, C, H, W))
    return (gx, gf) if want_flow_grad else gx


def sobel_forward(x, ksize=5, prec="f32"):
    x = _c(x, prec)
    N, _, H, W = x.shape
    out = np.empty((N, 2, H, W), _dt(prec))
    lib = _lib(prec)
    _fan(N, lambda i: lib.orc_sobel_forward(_p(x[i]), _p(out[i]), 1, H, W, int(ksize)))
    return out


def sobel_backward(grad_out, ksize=5, prec="f32"):
    grad_out = _c(grad_out, prec)
    N, _, H, W = grad_out.shape
    g = np.empty((N, 1, H, W), _dt(prec))
    lib = _lib(prec)
    _fan(N, lambda i: lib.orc_sobel_backward(_p(grad_out[i]), _p(g[i]), 1, H, W, int(ksize)))
    return g


def smooth_loss(disp, im, want_grad=True, prec="f32"):
    """-> (val, grad_disp or None); mean over the whole batch (model/networks.py:431)."""
    disp, im = _c(disp, prec), _c(im, prec)
    N, _, H, W = disp.shape
    g = np.empty_like(disp) if want_grad else None
    lib = _lib(prec)
    vals = _fan(N, lambda i: lib.orc_smooth_loss(_p(disp[i]), _p(im[i]), _p(g[i]) if want_grad else None, 1, H, W))
    if want_grad:
        g /= N
    return float(np.mean(vals)), g


def pattern_loss(disp, im, std, pattern, block_size=9, type=3, eps=0.5, want_grad=True, prec="f32"):
    """RectifiedPatternSimilarityLoss (model/networks.py:354-377).

    -> dict(val, num, den, diff, proj, grad_disp) ; the ratio is formed over the whole batch.
    """
    disp, im, std, pattern = _c(disp, prec), _c(im, prec), _c(std, prec), _c(pattern, prec)
    N, _, H, W = disp.shape
    pat = pattern.reshape(H, W)
    proj, diff = np.empty_like(disp), np.empty_like(disp)
    g = np.empty_like(disp) if want_grad else None
    nums = (ctypes.c_double * N)()
    dens = (ctypes.c_double * N)()
    lib = _lib(prec)

    def one(i):
        return lib.orc_pattern_loss(_p(disp[i]), _p(im[i]), _p(std[i]) if std is not None else None, _p(pat),
                                    _p(proj[i]), _p(diff[i]), _p(g[i]) if want_grad else None,
                                    ctypes.byref(nums, 8 * i), ctypes.byref(dens, 8 * i),
                                    1, H, W, int(block_size), int(type), _real(prec, eps))
    if any(_fan(N, one)):
        raise Exception("invalid loss type")
    nums, dens = np.array(nums[:]), np.array(dens[:])
    num, den = float(nums.sum()), float(dens.sum())
    if want_grad:
        # per-frame call normalised by the frame's own den; rescale to the batch-wide den
        g *= (dens / den).reshape(N, 1, 1, 1).astype(g.dtype)
    return dict(val=num / den, num=num, den=den, diff=diff, proj=proj, grad_disp=g)


def flow_consistency_dir(depth0, depth1, R0, t0, R1, t1, flow0, flow1, amb0, amb1, K, ray, clamp=-1.0,
                         primary_depth1=None, fb_scale=0.02, want_grad=True, prec="f32"):
    """One direction of the flow-consistency loss (model/networks.py:619-655 / 564-601).
    -> dict(loss, num, den, mask, orig_mask, grad_depth0, grad_depth1)  (gradients of loss = num/(den+1e-8))"""
    c = lambda a: _c(a, prec)
    depth0, depth1, R0, t0, R1, t1 = c(depth0), c(depth1), c(R0), c(t0), c(R1), c(t1)
    flow0, flow1, amb0, amb1, K, ray = c(flow0), c(flow1), c(amb0), c(amb1), c(K), c(ray)
    pd1 = c(primary_depth1)
    bs, _, H, W = depth0.shape
    mask, om = np.empty_like(depth0), np.empty_like(depth0)
    g0 = np.empty_like(depth0) if want_grad else None
    g1 = np.empty_like(depth0) if want_grad else None
    num, den = ctypes.c_double(), ctypes.c_double()
    _lib(prec).orc_flow_consistency(_p(depth0), _p(depth1), _p(R0), _p(t0), _p(R1), _p(t1), _p(flow0), _p(flow1),
                                    _p(amb0), _p(amb1), int(amb0.shape[1]), _p(pd1), _p(K), _p(ray),
                                    _real(prec, clamp), _real(prec, fb_scale), _p(mask), _p(om), _p(g0), _p(g1),
                                    ctypes.byref(num), ctypes.byref(den), bs, H, W)
    return dict(loss=num.value / (den.value + 1e-8), num=num.value, den=den.value, mask=mask, orig_mask=om,
                grad_depth0=g0, grad_depth1=g1)


def make_rays(K, H, W):
    """ray = [u, v, 1] @ Ki^T computed in float64 and rounded to float32 (model/networks.py:443-449)."""
    u, v = np.meshgrid(range(W), range(H))
    uv = np.stack((u, v, np.ones_like(u)), axis=2).reshape(-1, 3)
    return (uv @ np.linalg.inv(np.asarray(K, np.float64)).T).astype(np.float32)
