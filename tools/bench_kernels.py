"""Per-kernel timings (CUDA events, device-resident inputs larger than L2) for the secondary shapes of the path:
the DIS-MF flow warps, mse/sad window losses, LCN, smoothness, flow-consistency.  Prints one JSON line per kernel
with algorithmic GB/s against the measured HBM peak.  Run on the GPU box:  python tools/bench_kernels.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depthinspace_b200 import _ops, synth  # noqa: E402

PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, algo_bytes, **kw):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel=name, ms=round(ms, 4), algorithmic_GBps=round(gbs, 1), frac_of_measured_hbm=round(gbs / PEAK, 4), **kw)), flush=True)


def main():
    dev = torch.device("cuda")
    H, W = synth.DATASET_HW
    P = H * W
    N = 256
    x = torch.rand(N, 1, H, W, device=dev)
    report("lcn_forward r5", timeit(lambda: _ops.lcn_forward(x, 5, 0.05)), 12 * P * N, frames=N)
    disp = torch.rand(N, 1, H, W, device=dev) * 60
    amb = torch.rand(N, 1, H, W, device=dev)
    report("smooth_loss fwd+grad", timeit(lambda: _ops.smooth_loss_forward(disp, amb, True)), 12 * P * N, frames=N)
    report("smooth_loss fwd only", timeit(lambda: _ops.smooth_loss_forward(disp, amb, False)), 8 * P * N, frames=N)
    pat = torch.rand(1, 1, H, W, device=dev)
    im, std = _ops.lcn_forward(x, 5, 0.05)
    for t in ("mse", "sad", "census_mse", "census_sad"):
        report(f"pattern_loss single-scale {t} k9 fwd+grad", timeit(lambda: _ops.pattern_loss_forward(disp, im, std, pat, 9, t, 0.5, False, False, True)), 16 * P * N, frames=N)
    report("pattern_loss single-scale census_sad k9 fwd only", timeit(lambda: _ops.pattern_loss_forward(disp, im, std, pat, 9, 3, 0.5, False, False, False)), 12 * P * N, frames=N)
    d4 = [torch.rand(N, 1, H, W, device=dev) * 60 for _ in range(4)]
    report("pattern_loss 4-scale census_sad k9 fwd+grad", timeit(lambda: _ops.pattern_loss_multi_forward(d4, im, std, pat, 9, 3, 0.5, True)), 40 * P * N, frames=N)
    for t in ("mse", "sad"):   # point-wise kernel behind one box filter of the weights: im, std, M once + (disp, grad) per scale
        report(f"pattern_loss 4-scale {t} k9 fwd+grad", timeit(lambda: _ops.pattern_loss_multi_forward(d4, im, std, pat, 9, t, 0.5, True)), 40 * P * N, frames=N)
    report("pattern_loss single-scale mse k9 fwd+grad+map (tile kernel)", timeit(lambda: _ops.pattern_loss_forward(disp, im, std, pat, 9, "mse", 0.5, False, True, True)), 20 * P * N, frames=N)
    g = torch.rand(N, 1, H, W, device=dev)
    one, den = torch.ones(1, device=dev), torch.full((1,), 3.0, device=dev)
    report("scale_by_device_scalar", timeit(lambda: _ops.scale_by_device_scalar(g, one, den)), 8 * P * N, frames=N)
    es, ta = torch.randn(N, 1, H, W, device=dev), torch.randn(N, 1, H, W, device=dev)
    for t in ("mse", "census_sad"):
        report(f"photometric_loss_forward {t} k9", timeit(lambda: _ops.photometric_loss_forward(es, ta, 9, t, 0.5)), 12 * P * N, frames=N)
        report(f"photometric_loss_backward {t} k9", timeit(lambda: _ops.photometric_loss_backward(es, ta, g, 9, t, 0.5)), 16 * P * N, frames=N)
    # DIS-MF fusion warps: [bs, 32, 256, 216] and [bs, 32, 128, 108]
    for (h, w, bs) in ((256, 216, 32), (128, 108, 32)):
        p = h * w
        f = torch.from_numpy(synth.make_flows(bs, (h, w), max_mag=6.0)[0]).to(dev)
        feat = torch.randn(bs, 32, h, w, device=dev)
        go = torch.randn_like(feat)
        report(f"flow_warp_forward C=32 {h}x{w}", timeit(lambda: _ops.flow_warp_forward(feat, f)), (8 + 8 * 32) * p * bs, samples=bs)
        report(f"flow_warp_backward(x) C=32 {h}x{w}", timeit(lambda: _ops.flow_warp_backward(None, f, go, True, False)), (8 + 8 * 32) * p * bs, samples=bs)
    # FuseNet gather (own frame + 3 warped neighbours per target frame) and the Conv3D neighbour selection behind it
    tl, bs, C, h, w = 4, 32, 32, 256, 216
    p = h * w
    feat5 = torch.randn(tl, bs, C, h, w, device=dev)
    fl = {(i, j): torch.from_numpy(synth.make_flows(bs, (h, w), max_mag=6.0, seed=3 * i + j)[0]).to(dev)
          for i in range(tl) for j in range(tl) if i != j}
    go6 = torch.randn(tl, tl, bs, C, h, w, device=dev)
    report("flow_warp_gather_all_forward tl=4 C=32 256x216", timeit(lambda: _ops.flow_warp_gather_all_forward(feat5, fl)),
           4 * C * p * bs * (tl + tl * tl), samples=bs)
    report("flow_warp_gather_all_backward tl=4 C=32 256x216", timeit(lambda: _ops.flow_warp_gather_all_backward(fl, go6)),
           4 * C * p * bs * (tl + tl * tl), samples=bs)
    del go6
    bs = 4
    xyz = torch.randn(tl, bs, 3, h, w, device=dev)
    feat4 = torch.randn(tl, bs, C, h, w, device=dev)
    mask = (torch.rand(tl, bs, 1, h, w, device=dev) > 0.2).float()
    xyz_nb, feat_nb, idx, _ = _ops.conv3d_gather_forward(xyz, feat4, mask, 3, 1, 9)
    nb_bytes = 4 * (3 + C) * 9 * p * bs
    report("conv3d_gather_forward tl=4 C=32 256x216 (top-9 of 36)", timeit(lambda: _ops.conv3d_gather_forward(xyz, feat4, mask, 3, 1, 9)),
           4 * (3 + C + 1) * tl * p * bs + nb_bytes + 9 * p * bs, samples=bs)
    g_nb = torch.randn_like(feat_nb)
    report("conv3d_gather_backward(feat) tl=4 C=32 256x216", timeit(lambda: _ops.conv3d_gather_backward(None, g_nb, idx, (tl, bs, C, h, w), 3, 1, 9, False, True)),
           4 * C * 9 * p * bs + 9 * p * bs + 4 * C * tl * p * bs, samples=bs)
    # flow-consistency loss, one pair, both directions
    bs = 64
    gm = synth.make_geometry(2, (H, W), seed=1)
    rep = lambda k: torch.from_numpy(np.concatenate([gm[k]] * (bs // 2))).to(dev)
    from depthinspace_b200 import networks
    K = torch.from_numpy(gm["K"].astype(np.float64))
    mod = networks.Single_Frame_Flow_Consistency_Loss(K, torch.linalg.inv(K), H, W, clamp=0.1)
    args = [rep(k) for k in ("depth0", "depth1", "R0", "t0", "R1", "t1", "flow01", "flow10", "amb0", "amb1")]
    args[0].requires_grad_(True); args[1].requires_grad_(True)

    def fc():
        loss = mod(*args)[0]
        loss.backward()
    report("flow_consistency pair fwd+bwd (2 dirs)", timeit(fc), 2 * (6 * 4 + 2 * 4 + 2 * 4) * P * bs, samples=bs)


if __name__ == "__main__":
    main()
