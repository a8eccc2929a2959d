"""Generate the committed golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden

Every case stores the inputs, the reference's fp32 outputs/gradients and the outputs of the same
reference module run in fp64 (`*_f64`, the tie-breaker).  The reference has no tests or golden
vectors of its own, so these files are the pin for the oracle and, through it,
for the CUDA path.  Arithmetic caveat: generated on CPU -- torch's CPU grid_sample / scalar
division round differently from the CUDA kernels the reference runs in production, so sampled
values carry ~1e-6 of coordinate noise (bilinear interpolation is continuous, so no more).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depthinspace_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TYPES = ("mse", "sad", "census_mse", "census_sad")


def _np(t):
    return t.detach().cpu().numpy()


def case_lcn(ref, rng):
    out = {}
    for name, hw, radius in (("a", (24, 36), 5), ("b", (17, 21), 3)):
        x = rng.random((2, 1) + hw).astype(np.float32)
        x[0, 0, : hw[0] // 2] = 0.25  # flat region: worst case for E[x^2]-mu^2 cancellation
        mod = ref.networks.LCN(radius, 0.05)
        with torch.no_grad():
            l32, s32 = mod(torch.from_numpy(x))
            l64, s64 = mod.double()(torch.from_numpy(x).double())
        out.update({f"{name}_x": x, f"{name}_radius": radius, f"{name}_lcn": _np(l32), f"{name}_std": _np(s32),
                    f"{name}_lcn_f64": _np(l64), f"{name}_std_f64": _np(s64)})
    np.savez_compressed(os.path.join(OUT, "lcn.npz"), **out)


def case_photometric(ref, rng):
    out = {}
    for name, shape, k, eps in (("k9", (2, 1, 20, 28), 9, 0.5), ("k5c2", (1, 2, 13, 11), 5, 0.1), ("k3", (1, 1, 6, 7), 3, 0.5)):
        es = rng.standard_normal(shape).astype(np.float32)
        ta = rng.standard_normal(shape).astype(np.float32)
        go = rng.random((shape[0], 1) + shape[2:]).astype(np.float32)
        out.update({f"{name}_es": es, f"{name}_ta": ta, f"{name}_go": go, f"{name}_k": k, f"{name}_eps": eps})
        for t in TYPES:
            for dt, suf in ((torch.float32, ""), (torch.float64, "_f64")):
                e = torch.from_numpy(es).to(dt).requires_grad_(True)
                o = ref.ext_functions.photometric_loss_pytorch(e, torch.from_numpy(ta).to(dt), k, t, eps)
                o.backward(torch.from_numpy(go).to(dt))
                out[f"{name}_{t}_out{suf}"] = _np(o)
                out[f"{name}_{t}_grad{suf}"] = _np(e.grad)
    np.savez_compressed(os.path.join(OUT, "photometric.npz"), **out)


def case_pattern_loss(ref, rng):
    hw = (40, 56)
    d = synth.make_frames(2, hw, "kinect", n_scales=1, max_disp=32, seed=7)
    out = {}
    for dt, suf in ((torch.float32, ""), (torch.float64, "_f64")):
        lcn = ref.networks.LCN(5, 0.05).to(dt)
        im = torch.from_numpy(d["im"]).to(dt)
        with torch.no_grad():
            im_l, im_s = lcn(im)
            pat_l, _ = lcn(torch.from_numpy(d["pattern"]).to(dt))
        disp = torch.from_numpy(d["disp_pred"][0]).to(dt)
        disp[0, 0, 3, :10] = 0.0      # integer source coordinates
        disp[0, 0, 5, :] = 100.0      # far left of the pattern: border clip, zero gradient
        disp[1, 0, 7, :] = -100.0     # far right
        if suf == "":
            out.update(disp=_np(disp), im_lcn=_np(im_l), im_std=_np(im_s), pattern_lcn=_np(pat_l))
        for lt in ("census_sad", "mse"):
            for use_std in (True, False):
                dd = disp.clone().requires_grad_(True)
                mod = ref.networks.RectifiedPatternSimilarityLoss(hw[0], hw[1], torch.cat([pat_l] * 3, 1), loss_type=lt)
                mod.uv0 = mod.uv0.to(dt)
                val, proj = mod(dd, im_l, im_s if use_std else None)
                val.backward()
                key = f"{lt}_{'std' if use_std else 'nostd'}"
                out[f"{key}_val{suf}"] = _np(val)
                out[f"{key}_grad{suf}"] = _np(dd.grad)
                out[f"proj{suf}"] = _np(proj)
                out[f"pattern_mean{suf}"] = _np(mod.pattern)
        mod = ref.networks.RectifiedPatternSimilarityLoss(hw[0], hw[1], torch.cat([pat_l] * 3, 1))
        mod.uv0 = mod.uv0.to(dt)
        diff, _ = mod(disp, im_l, im_s, output_mean=False)
        out[f"census_sad_map{suf}"] = _np(diff)
    np.savez_compressed(os.path.join(OUT, "pattern_loss.npz"), **out)


def case_smooth(ref, rng):
    hw = (32, 40)
    d = synth.make_frames(2, hw, "default", n_scales=1, max_disp=32, seed=11)
    # keep the disparity free of exactly-flat patches: |g| has an ill-defined subgradient at g == 0
    disp0 = (d["disp_gt"] + rng.standard_normal(d["disp_gt"].shape).astype(np.float32)).astype(np.float32)
    out = dict(disp=disp0, ambient=d["ambient"])
    for dt, suf in ((torch.float32, ""), (torch.float64, "_f64")):
        disp = torch.from_numpy(disp0).to(dt).requires_grad_(True)
        amb = torch.from_numpy(d["ambient"]).to(dt)
        val = ref.networks.DisparitySmoothLoss().to(dt)(disp, amb)
        val.backward()
        sob = ref.networks.SobelFilter().to(dt)
        with torch.no_grad():
            g = sob(disp.detach())
        out.update({f"val{suf}": _np(val), f"grad{suf}": _np(disp.grad), f"sobel{suf}": _np(g)})
    np.savez_compressed(os.path.join(OUT, "smooth.npz"), **out)


def case_flow_warp(ref, rng):
    hw = (24, 32)
    x = rng.standard_normal((2, 3) + hw).astype(np.float32)
    f01, f10 = synth.make_flows(2, hw, max_mag=6.0, seed=5)
    f01[0, :, :3, :3] = 0.0        # integer coordinates
    f01[1, 0, 10, :] = 1000.0      # fully out of range -> zeros padding
    f01[1, 1, 12, :8] = -5.5
    go = rng.standard_normal(x.shape).astype(np.float32)
    out = dict(x=x, flow=f01, flow_back=f10, go=go)
    for dt, suf in ((torch.float32, ""), (torch.float64, "_f64")):
        xt = torch.from_numpy(x).to(dt).requires_grad_(True)
        ft = torch.from_numpy(f01).to(dt).requires_grad_(True)
        y = ref.multi_frame_networks.warp(xt, ft)
        y.backward(torch.from_numpy(go).to(dt))
        with torch.no_grad():  # multi_frame_networks.py:202-207
            f0 = torch.from_numpy(f01).to(dt)
            f10w = ref.multi_frame_networks.warp(torch.from_numpy(f10).to(dt), f0)
            mask = ((f0 + f10w) ** 2).sum(dim=1) < 0.5 + 0.01 * ((f0 ** 2).sum(dim=1) + (f10w ** 2).sum(dim=1))
        out.update({f"out{suf}": _np(y), f"grad_x{suf}": _np(xt.grad), f"grad_flow{suf}": _np(ft.grad),
                    f"flow_back_warped{suf}": _np(f10w), f"fb_mask{suf}": _np(mask.float().unsqueeze(1))})
    np.savez_compressed(os.path.join(OUT, "flow_warp.npz"), **out)


def case_flow_consistency(ref, rng):
    hw = (32, 44)
    g = synth.make_geometry(2, hw, seed=3)
    out = {k: v for k, v in g.items()}
    pd0, pd1 = (g["depth0"] + 0.002).astype(np.float32), (g["depth1"] - 0.002).astype(np.float32)
    out.update(primary_depth0=pd0, primary_depth1=pd1)
    for dt, suf in ((torch.float32, ""), (torch.float64, "_f64")):
        K = torch.from_numpy(g["K"].astype(np.float64)).to(dt)
        Ki = torch.from_numpy(np.linalg.inv(g["K"].astype(np.float64))).to(dt)
        T = lambda k: torch.from_numpy(g[k]).to(dt)
        for mf in (False, True):
            cls = ref.networks.Multi_Frame_Flow_Consistency_Loss if mf else ref.networks.Single_Frame_Flow_Consistency_Loss
            mod = cls(K, Ki, hw[0], hw[1], clamp=0.1)
            mod.ray, mod.u, mod.v = mod.ray.to(dt), mod.u.to(dt), mod.v.to(dt)
            d0, d1 = T("depth0").requires_grad_(True), T("depth1").requires_grad_(True)
            args = [d0, d1, T("R0"), T("t0"), T("R1"), T("t1"), T("flow01"), T("flow10"), T("amb0"), T("amb1")]
            key = "mf" if mf else "sf"
            if mf:
                loss = mod(*args, torch.from_numpy(pd0).to(dt), torch.from_numpy(pd1).to(dt))
            else:
                loss, m0, m1, om = mod(*args)
                out.update({f"sf_mask0{suf}": _np(m0), f"sf_mask1{suf}": _np(m1), f"sf_orig_mask{suf}": np.asarray(om)})
            loss.backward()
            out.update({f"{key}_loss{suf}": _np(loss), f"{key}_grad0{suf}": _np(d0.grad), f"{key}_grad1{suf}": _np(d1.grad)})
    np.savez_compressed(os.path.join(OUT, "flow_consistency.npz"), **out)


def main():
    ref = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(42)
    rng = np.random.default_rng(42)
    for fn in (case_lcn, case_photometric, case_pattern_loss, case_smooth, case_flow_warp, case_flow_consistency):
        fn(ref, rng)
        print("wrote", fn.__name__)


if __name__ == "__main__":
    main()
