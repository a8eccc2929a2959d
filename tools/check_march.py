"""A/B check of the pair-symmetric marching kernel (csrc/pattern_march.cuh) against the round-1 tile kernel
(DIS_MULTI_IMPL=tile) and the fp64 C oracle, over shapes / windows / band plans, plus timings at the bench size.

    python tools/check_march.py [--time]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from depthinspace_b200 import _ops, synth  # noqa: E402
from oracle import c_oracle  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run(disps, im, std, pat, k, lt, impl, env=None):
    old = {}
    env = dict(env or {})
    env["DIS_MULTI_IMPL"] = impl
    for key, v in env.items():
        old[key] = os.environ.get(key)
        os.environ[key] = str(v)
    try:
        out3, grads = _ops.pattern_loss_multi_forward(disps, im, std, pat, k, lt, 0.5, True)
        torch.cuda.synchronize()
    finally:
        for key, v in old.items():
            if v is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = v
    return out3.cpu().numpy().astype(np.float64), [g.cpu().numpy().astype(np.float64) for g in grads]


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def frac(a, b, tol):
    return float((np.abs(a - b) > tol * np.abs(b).max()).mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--small", action="store_true", help="skip the large shapes (for compute-sanitizer runs)")
    a = ap.parse_args()
    worst = {"val_tile": 0.0, "grad_tile": 0.0, "val_o64": 0.0, "grad_o64_frac": 0.0}
    ok = True
    cases = [
        # (N, (H, W), k, S, lt, env)
        (2, (70, 150), 9, 4, "census_sad", {}),
        (2, (70, 150), 9, 4, "census_sad", {"DIS_MARCH_BAND_ROWS": 20}),
        (2, (70, 150), 9, 4, "census_sad", {"DIS_MARCH_BAND_ROWS": 16, "DIS_MARCH_WARPS": 2}),
        (2, (70, 150), 9, 2, "census_mse", {"DIS_MARCH_BAND_ROWS": 24, "DIS_MARCH_WARPS": 3}),
        (3, (33, 47), 5, 4, "census_sad", {}),
        (3, (33, 47), 5, 2, "census_sad", {"DIS_MARCH_BAND_ROWS": 8, "DIS_MARCH_WARPS": 1}),
        (2, (64, 96), 13, 4, "census_sad", {}),
        (2, (64, 96), 13, 4, "census_mse", {"DIS_MARCH_BAND_ROWS": 28, "DIS_MARCH_WARPS": 2}),
        (2, (40, 70), 15, 4, "census_sad", {}),
        (2, (40, 70), 15, 2, "census_sad", {"DIS_MARCH_BAND_ROWS": 32, "DIS_MARCH_WARPS": 2}),
        (2, (33, 65), 3, 4, "census_sad", {"DIS_MARCH_BAND_ROWS": 8}),
        (2, (33, 65), 7, 4, "census_sad", {"DIS_MARCH_BAND_ROWS": 12, "DIS_MARCH_WARPS": 1}),
        (2, (2, 2), 9, 4, "census_sad", {}),
        (1, (5, 3), 9, 2, "census_mse", {}),
        (1, (9, 300), 11, 4, "census_sad", {}),
        (1, (300, 9), 11, 4, "census_sad", {}),
        (2, (128, 216), 9, 4, "census_sad", {}),
        (2, (512, 432), 9, 4, "census_sad", {}),
        (1, (480, 640), 9, 4, "census_sad", {}),
    ]
    for (N, hw, k, S, lt, env) in cases:
        if a.small and hw[0] * hw[1] > 100 * 160:
            continue
        if min(hw) >= 24:
            d = synth.make_frames(N, hw, "kinect", n_scales=S, max_disp=min(48, max(2, hw[1] // 2)), seed=k + S)
            lcn_im, std = _ops.lcn_forward(dev(d["im"]), 5, 0.05)
            pat = _ops.lcn_forward(dev(d["pattern"]), 5, 0.05)[0].reshape(hw)
        else:   # tiny / degenerate shapes: plain random planes
            rng = np.random.default_rng(k + S)
            d = {"disp_pred": [(rng.random((N, 1) + hw) * hw[1]).astype(np.float32) for _ in range(S)]}
            lcn_im = dev(rng.standard_normal((N, 1) + hw).astype(np.float32))
            std = dev((rng.random((N, 1) + hw) + 0.05).astype(np.float32))
            pat = dev(rng.standard_normal(hw).astype(np.float32))
        disps = [dev(p) for p in d["disp_pred"]]
        o_m, g_m = run(disps, lcn_im, std, pat, k, lt, "march", env)
        o_t, g_t = run(disps, lcn_im, std, pat, k, lt, "tile")
        o_m2, g_m2 = run(disps, lcn_im, std, pat, k, lt, "march", env)
        repro = np.array_equal(o_m, o_m2) and all(np.array_equal(x, y) for x, y in zip(g_m, g_m2))
        rv = rel(o_m[:, :2], o_t[:, :2])
        rg = max(rel(x, y) for x, y in zip(g_m, g_t))
        line = {"N": N, "hw": hw, "k": k, "S": S, "lt": lt, "env": env, "val_vs_tile": rv, "grad_vs_tile": rg, "repro": repro}
        if not a.no_oracle and hw[0] * hw[1] <= 130 * 220:
            tid = c_oracle.TYPES[lt]
            rvo, fro, rgo = 0.0, 0.0, 0.0
            for s in range(S):
                o64 = c_oracle.pattern_loss(d["disp_pred"][s], lcn_im.cpu().numpy(), std.cpu().numpy(), pat.cpu().numpy(), k, tid, 0.5, True, "f64")
                rvo = max(rvo, abs(o_m[s, 0] - o64["num"]) / max(abs(o64["num"]), 1e-30), abs(o_m[s, 1] - o64["den"]) / o64["den"])
                gref = o64["grad_disp"] * o64["den"]      # un-normalised gradient of the numerator
                fro = max(fro, frac(g_m[s], gref, 1e-5))
                rgo = max(rgo, rel(g_m[s], gref))
            line.update(val_vs_o64=rvo, grad_vs_o64_max=rgo, grad_vs_o64_outlier_frac=fro)
            ok &= rvo < 1e-5 and fro < 1e-2
        ok &= repro and rv < 5e-6 and (rg < 1e-4)
        print(json.dumps(line), flush=True)
    print("CHECK", "OK" if ok else "FAILED", flush=True)
    if a.time:
        N, hw = 256, (512, 432)
        d = synth.make_frames(8, hw, "default", n_scales=4, max_disp=128, seed=0)
        rep = N // 8
        im = dev(np.tile(d["im"], (rep, 1, 1, 1)))
        lcn_im, std = _ops.lcn_forward(im, 5, 0.05)
        pat = _ops.lcn_forward(dev(d["pattern"]), 5, 0.05)[0].reshape(hw)
        disps = [dev(np.tile(p, (rep, 1, 1, 1))) for p in d["disp_pred"]]
        configs = [("tile", {})] + [("march", {"DIS_MARCH_BAND_ROWS": r}) for r in (0, 44, 66, 88, 104, 130, 174, 260, 520)]
        for impl, env in configs:
            for key, v in env.items():
                if v:
                    os.environ[key] = str(v)
            os.environ["DIS_MULTI_IMPL"] = impl
            for _ in range(3):
                _ops.pattern_loss_multi_forward(disps, lcn_im, std, pat, 9, "census_sad", 0.5, True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(10):
                _ops.pattern_loss_multi_forward(disps, lcn_im, std, pat, 9, "census_sad", 0.5, True)
            ev[1].record()
            torch.cuda.synchronize()
            print(json.dumps({"impl": impl, "env": env, "ms": ev[0].elapsed_time(ev[1]) / 10}), flush=True)
            for key in env:
                os.environ.pop(key, None)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
