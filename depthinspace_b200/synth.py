"""Synthetic structured-light frames of the dataset's shape (no dataset / network here).

Mimics what the reference's renderer produces (data/create_syn_data.py:106-144, 227, 295:
a board plus a few objects; IR = 0.6 * warped pattern + 0.4 * ambient; sensor noise as in
data/data_manipulation.py:170-192).  At the dataset shape (512x432) the projector patterns are the
REFERENCE's own default / kinect / real patterns after its read_pattern_file + remap + post_process
pipeline (tests/golden/patterns_512x432.npz, written by oracle/gen_patterns.py from the reference's
PNGs; intrinsics K and baseline of each pattern are stored beside them).  For any other shape (the
reference defines its patterns for 512x432 only) or if the fixture is missing, a procedural random
dot field with the same mean intensity (0.037 / 0.183 / 0.312) stands in.

Everything is plain numpy on the host, seeded, and shape-parametric.
"""
import os

import numpy as np

PATTERN_DENSITY = {"default": 0.037, "kinect": 0.183, "real": 0.312}
DATASET_HW = (512, 432)      # data/create_syn_data.py:301-302
CTD_HW = (480, 640)          # literal shape in BASELINE.json configs[0]


def _blur3(a):
    p = np.pad(a, 1, mode="edge")
    return (p[:-2, :-2] + p[:-2, 2:] + p[2:, :-2] + p[2:, 2:] + 2 * (p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:])
            + 4 * p[1:-1, 1:-1]) / 16.0


_PATTERN_FILE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                             "patterns_512x432.npz")
_PATTERN_CACHE = {}


def reference_pattern(kind):
    """-> (pattern [512,432] float32, K [3,3], baseline) of the reference, or None when the fixture is absent."""
    if not _PATTERN_CACHE and os.path.exists(_PATTERN_FILE):
        with np.load(_PATTERN_FILE) as z:
            for k in PATTERN_DENSITY:
                _PATTERN_CACHE[k] = (z[k].astype(np.float32), z[k + "_K"].astype(np.float32), float(z[k + "_baseline"]))
    return _PATTERN_CACHE.get(kind)


def pattern_source(kind="default", hw=DATASET_HW):
    """'reference' when dot_pattern() returns the reference's remapped pattern, 'procedural' for the stand-in."""
    return "reference" if tuple(hw) == DATASET_HW and reference_pattern(kind) is not None else "procedural"


def dot_pattern(kind="default", hw=DATASET_HW, seed=42):
    """Projector dot pattern [H,W] float32 in [0,1]: the reference's at 512x432, a procedural stand-in otherwise."""
    if pattern_source(kind, hw) == "reference":
        return reference_pattern(kind)[0].copy()
    rng = np.random.default_rng(seed + sum(map(ord, kind)))
    H, W = hw
    density = PATTERN_DENSITY[kind]
    dots = (rng.random((H, W)) < density * 0.45).astype(np.float32)
    pat = _blur3(dots) * 4.0
    if kind == "real":  # the real projector image is soft and has a bright pedestal
        pat = 0.5 * _blur3(pat) + 0.18
    pat = np.clip(pat, 0.0, 1.0)
    pat *= density / max(float(pat.mean()), 1e-6)
    return np.clip(pat, 0.0, 1.0).astype(np.float32)


def _smooth_field(rng, hw, cells=6):
    """Low-frequency field in [0,1] by bilinear upsampling of a coarse random grid."""
    H, W = hw
    g = rng.random((cells + 1, cells + 1)).astype(np.float32)
    ys = np.linspace(0, cells, H, dtype=np.float32)
    xs = np.linspace(0, cells, W, dtype=np.float32)
    y0 = np.clip(np.floor(ys).astype(int), 0, cells - 1)
    x0 = np.clip(np.floor(xs).astype(int), 0, cells - 1)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx).astype(np.float32)


def gt_disparity(rng, hw, lo=2.0, hi=60.0):
    """Board (sum of planes) + 4 rectangular objects, range [lo, hi] px."""
    H, W = hw
    v, u = np.meshgrid(np.linspace(0, 1, H, dtype=np.float32), np.linspace(0, 1, W, dtype=np.float32), indexing="ij")
    d = np.zeros(hw, np.float32)
    for _ in range(3):
        a, b, c = rng.uniform(-1, 1, 3)
        d += a * u + b * v + c
    d = (d - d.min()) / max(float(d.max() - d.min()), 1e-6)
    d = lo + (0.25 + 0.35 * d) * (hi - lo)
    for _ in range(4):
        h0, w0 = rng.integers(0, H - H // 4), rng.integers(0, W - W // 4)
        hh, ww = rng.integers(H // 8, H // 3), rng.integers(W // 8, W // 3)
        d[h0:h0 + hh, w0:w0 + ww] += rng.uniform(3.0, 15.0)
    return np.clip(d, lo, hi).astype(np.float32)


def _warp_rows(pattern, disp):
    """Horizontal bilinear warp (border clamp) used only to render plausible IR frames."""
    H, W = pattern.shape
    x = np.clip(np.arange(W, dtype=np.float32)[None, :] - disp, 0, W - 1)
    x0 = np.clip(np.floor(x).astype(int), 0, W - 2)
    f = x - x0
    rows = np.arange(H)[:, None]
    return pattern[rows, x0] * (1 - f) + pattern[rows, x0 + 1] * f


def make_frames(n_frames, hw=DATASET_HW, pattern_kind="default", n_scales=4, max_disp=128.0, seed=42):
    """-> dict of float32 arrays: pattern [1,1,H,W]; im, ambient, disp_gt [N,1,H,W];
    disp_pred: list of n_scales [N,1,H,W] (network-output stand-ins, scale s noisier)."""
    rng = np.random.default_rng(seed)
    H, W = hw
    pat = dot_pattern(pattern_kind, hw, seed)
    im = np.empty((n_frames, 1, H, W), np.float32)
    amb = np.empty_like(im)
    dgt = np.empty_like(im)
    for i in range(n_frames):
        d = gt_disparity(rng, hw)
        a = 0.1 + 0.6 * _smooth_field(rng, hw)
        ir = 0.6 * _warp_rows(pat, d) + 0.4 * a
        sigma = (3.0 / 255.0) * rng.random()
        ir = ir + rng.normal(0, 1, hw).astype(np.float32) * sigma
        im[i, 0], amb[i, 0], dgt[i, 0] = np.clip(ir, 0, 1), a, d
    preds = []
    for s in range(n_scales):
        p = dgt + rng.normal(0, 1, dgt.shape).astype(np.float32) * np.float32(np.sqrt(2.0 ** s) * 0.5)
        preds.append(np.clip(p, 1e-3, max_disp / 2 ** s - 1e-3).astype(np.float32))
    return dict(pattern=pat[None, None], im=im, ambient=amb, disp_gt=dgt, disp_pred=preds)


def make_flows(n, hw, max_mag=8.0, seed=42, inconsistent=0.05):
    """Smooth forward flow, approximately inverse backward flow, 5 % inconsistent pixels."""
    rng = np.random.default_rng(seed + 7)
    H, W = hw
    f01 = np.empty((n, 2, H, W), np.float32)
    for i in range(n):
        f01[i, 0] = (2 * _smooth_field(rng, hw) - 1) * max_mag
        f01[i, 1] = (2 * _smooth_field(rng, hw) - 1) * max_mag
    f10 = -f01 + rng.normal(0, 0.05, f01.shape).astype(np.float32)
    bad = rng.random((n, 1, H, W)) < inconsistent
    f10 = np.where(bad, f10 + rng.normal(0, 3.0, f01.shape).astype(np.float32), f10).astype(np.float32)
    return f01, f10


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _bilinear_zero(img, x, y):
    H, W = img.shape
    x0, y0 = np.floor(x).astype(int), np.floor(y).astype(int)
    out = np.zeros_like(x, dtype=np.float64)
    for dy in (0, 1):
        for dx in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            w = (1 - np.abs(x - xi)) * (1 - np.abs(y - yi))
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            out += np.where(ok, img[np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)] * w, 0.0)
    return out


def make_geometry(bs, hw, seed=42, flow_noise=0.15, bump=0.02):
    """Two views of a slanted plane with small bumps: intrinsics, poses (row-vector convention of the
    reference: X_world = (X_cam - t) R), per-view depth, optical flow both ways (from the geometry plus
    noise, a few percent grossly wrong), ambient images related by the flow.  float32 arrays:
    K [3,3], R0,R1 [bs,3,3], t0,t1 [bs,3], depth0,depth1,amb0,amb1 [bs,1,H,W], flow01,flow10 [bs,2,H,W]."""
    rng = np.random.default_rng(seed + 101)
    H, W = hw
    f = 1.3 * W
    K = np.array([[f, 0, (W - 1) / 2.0], [0, f, (H - 1) / 2.0], [0, 0, 1]], np.float64)
    Ki = np.linalg.inv(K)
    v, u = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    ray = np.stack((u, v, np.ones_like(u)), -1).reshape(-1, 3) @ Ki.T
    out = {k: [] for k in ("R0", "t0", "R1", "t1", "depth0", "depth1", "flow01", "flow10", "amb0", "amb1")}

    def plane_depth(R, t, n, d):   # n . ((depth ray - t) R) = d
        return (d + (t @ R) @ n) / ((ray @ R) @ n)

    def project(depth, Ra, ta, Rb, tb):
        X = ((depth[:, None] * ray - ta) @ Ra) @ Rb.T + tb
        p = X @ K.T
        return p[:, 0] / p[:, 2], p[:, 1] / p[:, 2]

    for _ in range(bs):
        n = np.array([rng.uniform(-0.25, 0.25), rng.uniform(-0.25, 0.25), 1.0])
        d = rng.uniform(1.2, 2.0)
        R0, R1 = _rot(*rng.uniform(-0.03, 0.03, 3)), _rot(*rng.uniform(-0.03, 0.03, 3))
        t0, t1 = rng.uniform(-0.03, 0.03, 3), rng.uniform(-0.03, 0.03, 3)
        dep0, dep1 = plane_depth(R0, t0, n, d), plane_depth(R1, t1, n, d)
        u1, v1 = project(dep0, R0, t0, R1, t1)
        u0, v0 = project(dep1, R1, t1, R0, t0)
        f01 = np.stack((u1 - u.ravel(), v1 - v.ravel())).reshape(2, H, W)
        f10 = np.stack((u0 - u.ravel(), v0 - v.ravel())).reshape(2, H, W)
        f01 += rng.normal(0, flow_noise, f01.shape)
        f10 += rng.normal(0, flow_noise, f10.shape)
        bad = rng.random((1, H, W)) < 0.04
        f01 = np.where(bad, f01 + rng.normal(0, 4.0, f01.shape), f01)
        a0 = 0.1 + 0.6 * _smooth_field(rng, hw).astype(np.float64)
        a1 = _bilinear_zero(a0, u + f10[0], v + f10[1]) + rng.normal(0, 0.003, hw)
        dep0 = dep0.reshape(H, W) + bump * _smooth_field(rng, hw, cells=10)
        dep1 = dep1.reshape(H, W) + bump * _smooth_field(rng, hw, cells=10)
        for k, val in (("R0", R0), ("t0", t0), ("R1", R1), ("t1", t1), ("depth0", dep0[None]), ("depth1", dep1[None]),
                       ("flow01", f01), ("flow10", f10), ("amb0", a0[None]), ("amb1", a1[None])):
            out[k].append(val)
    res = {k: np.stack(v).astype(np.float32) for k, v in out.items()}
    res["K"] = K.astype(np.float32)
    return res
