"""The oracle (plain-C restatement, fp32 and fp64 builds, and the torch port) against the golden vectors
generated from the unmodified reference by oracle/gen_golden.py.  Runs everywhere (no reference, no GPU)."""
import numpy as np
import pytest
import torch

from helpers import assert_close, assert_scalar_close
from oracle import c_oracle, torch_port

TYPES = ("mse", "sad", "census_mse", "census_sad")
# fp32 oracle vs fp32 reference: two fp32 evaluations with different summation order
F32 = 2e-6
F64 = 1e-12


@pytest.mark.parametrize("case", ["a", "b"])
def test_lcn(golden, case):
    g = golden("lcn")
    x, r = g[f"{case}_x"], int(g[f"{case}_radius"])
    l64, s64 = c_oracle.lcn_forward(x, r, 0.05, "f64")
    assert_close(l64, g[f"{case}_lcn_f64"], F64, "lcn f64")
    assert_close(s64, g[f"{case}_std_f64"], F64, "std f64")
    l32, s32 = c_oracle.lcn_forward(x, r, 0.05, "f32")
    # E[x^2]-mu^2 cancels catastrophically on the flat half of the image: fp32 evaluations agree only to ~1e-3 in std
    assert_close(l32, g[f"{case}_lcn"], 1e-5, "lcn f32")
    assert_close(s32, g[f"{case}_std"], 5e-3, "std f32")
    lt, st = torch_port.lcn(torch.from_numpy(x), r, 0.05)
    assert_close(lt, g[f"{case}_lcn"], 1e-6, "lcn torch port")
    assert_close(st, g[f"{case}_std"], 1e-6, "std torch port")


@pytest.mark.parametrize("case", ["k9", "k5c2", "k3"])
@pytest.mark.parametrize("t", TYPES)
def test_photometric(golden, case, t):
    g = golden("photometric")
    es, ta, go = g[f"{case}_es"], g[f"{case}_ta"], g[f"{case}_go"]
    k, eps, tid = int(g[f"{case}_k"]), float(g[f"{case}_eps"]), c_oracle.TYPES[t]
    for prec, suf, tol in (("f64", "_f64", F64), ("f32", "", 5e-6)):
        out = c_oracle.photometric_forward(es, ta, k, tid, eps, prec)
        grad = c_oracle.photometric_backward(es, ta, go, k, tid, eps, prec)
        assert_close(out, g[f"{case}_{t}_out{suf}"], tol, f"fwd {prec}")
        assert_close(grad, g[f"{case}_{t}_grad{suf}"], tol, f"bwd {prec}")
    e = torch.from_numpy(es).requires_grad_(True)
    o = torch_port.photometric(e, torch.from_numpy(ta), k, t, eps)
    o.backward(torch.from_numpy(go))
    assert_close(o, g[f"{case}_{t}_out"], 1e-6, "fwd torch port")
    assert_close(e.grad, g[f"{case}_{t}_grad"], 1e-6, "bwd torch port")


def test_photometric_invalid_type():
    z = np.zeros((1, 1, 4, 4), np.float32)
    with pytest.raises(Exception, match="invalid loss type"):
        c_oracle.photometric_forward(z, z, 3, 4, 0.5)
    with pytest.raises(Exception, match="invalid loss type"):
        torch_port.photometric(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), 3, "ssim")


@pytest.mark.parametrize("lt", ["census_sad", "mse"])
@pytest.mark.parametrize("use_std", [True, False])
def test_pattern_loss(golden, lt, use_std):
    g = golden("pattern_loss")
    key = f"{lt}_{'std' if use_std else 'nostd'}"
    tid = c_oracle.TYPES[lt]
    std = g["im_std"] if use_std else None
    # fp64: same formulas, both sides essentially exact
    o = c_oracle.pattern_loss(g["disp"].astype(np.float64), g["im_lcn"], std, g["pattern_mean_f64"], 9, tid, 0.5, True, "f64")
    # inputs of the f64 golden run were produced by an fp64 LCN, so only the fp32-input run is comparable: use fp32 inputs
    o32 = c_oracle.pattern_loss(g["disp"], g["im_lcn"], std, g["pattern_mean"], 9, tid, 0.5, True, "f32")
    assert_scalar_close(o32["val"], g[f"{key}_val"], 2e-6, "val f32")
    # CPU torch and the CUDA-order oracle round the sampling coordinate differently (~1e-6 of a pixel)
    assert_close(o32["proj"], g["proj"], 2e-5, "pattern_proj")
    assert_close(o32["grad_disp"], g[f"{key}_grad"], 2e-5, "grad_disp")
    assert np.isfinite(o["val"])


def test_pattern_loss_map(golden):
    g = golden("pattern_loss")
    o32 = c_oracle.pattern_loss(g["disp"], g["im_lcn"], g["im_std"], g["pattern_mean"], 9, 3, 0.5, False, "f32")
    assert_close(o32["diff"], g["census_sad_map"], 2e-5, "per-pixel map")


def test_smooth(golden):
    g = golden("smooth")
    for prec, suf, tol in (("f64", "_f64", 1e-11), ("f32", "", 1e-5)):
        val, grad = c_oracle.smooth_loss(g["disp"], g["ambient"], True, prec)
        assert_scalar_close(val, g[f"val{suf}"], tol, f"val {prec}")
        assert_close(grad, g[f"grad{suf}"], tol, f"grad {prec}", outlier_frac=2e-3 if prec == "f32" else 0)
        assert_close(c_oracle.sobel_forward(g["disp"], 5, prec), g[f"sobel{suf}"], tol, f"sobel {prec}")
    d = torch.from_numpy(g["disp"]).requires_grad_(True)
    v = torch_port.smooth_loss(d, torch.from_numpy(g["ambient"]))
    v.backward()
    assert_scalar_close(v.item(), g["val"], 1e-6)
    assert_close(d.grad, g["grad"], 1e-6, "torch port grad", outlier_frac=2e-3)


def test_flow_warp(golden):
    g = golden("flow_warp")
    for prec, suf, tol in (("f64", "_f64", 1e-11), ("f32", "", 2e-5)):
        out, _, _ = c_oracle.flow_warp_forward(g["x"], g["flow"], prec)
        gx, gf = c_oracle.flow_warp_backward(g["x"], g["flow"], g["go"], True, prec)
        assert_close(out, g[f"out{suf}"], tol, f"out {prec}")
        assert_close(gx, g[f"grad_x{suf}"], tol, f"grad_x {prec}")
        assert_close(gf, g[f"grad_flow{suf}"], tol, f"grad_flow {prec}")
    y = torch_port.flow_warp(torch.from_numpy(g["x"]), torch.from_numpy(g["flow"]))
    assert_close(y, g["out"], 1e-6, "torch port")
    f10w = torch_port.flow_warp(torch.from_numpy(g["flow_back"]), torch.from_numpy(g["flow"]))
    m = torch_port.fb_mask(torch.from_numpy(g["flow"]), f10w)
    assert (m.numpy() != g["fb_mask"]).mean() < 1e-3


def _fc_dirs(g, mf, prec):
    ray = c_oracle.make_rays(g["K"], *g["depth0"].shape[-2:])
    kw = dict(clamp=-1.0 if mf else 0.1, prec=prec)
    A = c_oracle.flow_consistency_dir(g["depth0"], g["depth1"], g["R0"], g["t0"], g["R1"], g["t1"], g["flow01"], g["flow10"],
                                      g["amb0"], g["amb1"], g["K"], ray, primary_depth1=g["primary_depth1"] if mf else None, **kw)
    B = c_oracle.flow_consistency_dir(g["depth1"], g["depth0"], g["R1"], g["t1"], g["R0"], g["t0"], g["flow10"], g["flow01"],
                                      g["amb1"], g["amb0"], g["K"], ray, primary_depth1=g["primary_depth0"] if mf else None, **kw)
    return A, B


@pytest.mark.parametrize("mf", [False, True])
def test_flow_consistency(golden, mf):
    g = golden("flow_consistency")
    key = "mf" if mf else "sf"
    A, B = _fc_dirs(g, mf, "f64")
    assert_scalar_close(A["loss"] + B["loss"], float(g[f"{key}_loss_f64"]), 1e-10)
    assert_close(A["grad_depth0"] + B["grad_depth1"], g[f"{key}_grad0_f64"], 1e-9, "grad depth0")
    assert_close(A["grad_depth1"] + B["grad_depth0"], g[f"{key}_grad1_f64"], 1e-9, "grad depth1")
    if not mf:
        assert np.array_equal(A["mask"], g["sf_mask0_f64"]) and np.array_equal(B["mask"], g["sf_mask1_f64"])
        assert np.array_equal(A["orig_mask"][0, 0], g["sf_orig_mask_f64"])
    A32, B32 = _fc_dirs(g, mf, "f32")
    # fp32: thresholded masks may flip on a few pixels between the CPU-torch and the CUDA-order arithmetic
    assert_scalar_close(A32["loss"] + B32["loss"], float(g[f"{key}_loss"]), 1e-3)
    if not mf:
        assert (A32["mask"] != g["sf_mask0"]).mean() < 5e-3
