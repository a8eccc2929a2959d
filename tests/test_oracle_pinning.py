"""Pins the oracle against the UNMODIFIED reference imported from /root/reference (build container only;
skipped where the checkout is absent, e.g. on the GPU box -- there the committed golden vectors are the pin).

Parity status at the ext_cuda boundary: the reference's native kernel (Connecting-the-Dots torchext) is
un-vendored and unpinned; its in-tree definition photometric_loss_pytorch (model/ext_functions.py:156-183)
is what both the oracle and the CUDA path are held to."""
import numpy as np
import pytest
import torch

from helpers import assert_close, assert_scalar_close
from depthinspace_b200 import synth
from oracle import c_oracle, ref_shim, torch_port

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


@pytest.fixture(scope="module")
def frames(ref):
    hw = (48, 64)
    d = synth.make_frames(3, hw, "default", n_scales=2, max_disp=32, seed=3)
    lcn = ref.networks.LCN(5, 0.05)
    with torch.no_grad():
        im_l, im_s = lcn(torch.from_numpy(d["im"]))
        pat_l, _ = lcn(torch.from_numpy(d["pattern"]))
    return hw, d, im_l, im_s, pat_l


@pytest.mark.parametrize("t", ["mse", "sad", "census_mse", "census_sad"])
@pytest.mark.parametrize("k", [1, 3, 9])
def test_photometric_c_oracle_f64_equals_reference(ref, t, k):
    torch.manual_seed(k)
    es = torch.randn(2, 2, 14, 19, dtype=torch.float64, requires_grad=True)
    ta = torch.randn(2, 2, 14, 19, dtype=torch.float64)
    out = ref.ext_functions.photometric_loss_pytorch(es, ta, k, t, 0.3)
    go = torch.rand_like(out)
    out.backward(go)
    tid = c_oracle.TYPES[t]
    assert_close(c_oracle.photometric_forward(es.detach().numpy(), ta.numpy(), k, tid, 0.3, "f64"), out, 1e-13)
    assert_close(c_oracle.photometric_backward(es.detach().numpy(), ta.numpy(), go.numpy(), k, tid, 0.3, "f64"), es.grad, 1e-13)


def test_torch_port_is_bit_identical_on_cpu(ref, frames):
    hw, d, im_l, im_s, pat_l = frames
    a, b = torch_port.lcn(torch.from_numpy(d["im"]))
    assert torch.equal(a, im_l) and torch.equal(b, im_s)
    mod = ref.networks.RectifiedPatternSimilarityLoss(hw[0], hw[1], torch.cat([pat_l] * 3, 1))
    disp = torch.from_numpy(d["disp_pred"][0]).requires_grad_(True)
    v, p = mod(disp, im_l, im_s)
    v.backward()
    g_ref = disp.grad.clone()
    disp.grad = None
    v2, p2 = torch_port.pattern_loss(disp, im_l, im_s, mod.pattern)
    v2.backward()
    assert torch.equal(p, p2) and torch.equal(v, v2) and torch.equal(g_ref, disp.grad)
    amb = torch.from_numpy(d["ambient"])
    assert torch.equal(ref.networks.DisparitySmoothLoss()(disp, amb), torch_port.smooth_loss(disp, amb))
    assert torch.equal(ref.networks.SobelFilter()(amb), torch_port.sobel(amb))
    x = torch.randn(3, 4, *hw)
    fl = torch.from_numpy(synth.make_flows(3, hw)[0])
    assert torch.equal(ref.multi_frame_networks.warp(x, fl), torch_port.flow_warp(x, fl))
    for t in torch_port.LOSS_TYPES:
        assert torch.equal(ref.ext_functions.photometric_loss_pytorch(x, x.flip(0), 5, t, 0.2),
                           torch_port.photometric(x, x.flip(0), 5, t, 0.2))


def test_pattern_loss_c_oracle_f64_equals_reference_f64(ref, frames):
    hw, d, im_l, im_s, pat_l = frames
    mod = ref.networks.RectifiedPatternSimilarityLoss(hw[0], hw[1], torch.cat([pat_l] * 3, 1).double())
    mod.uv0 = mod.uv0.double()
    disp = torch.from_numpy(d["disp_pred"][1]).double()
    disp[0, 0, 2, :7] = 0.0
    disp[1, 0, 4, :] = 90.0
    disp.requires_grad_(True)
    v, p = mod(disp, im_l.double(), im_s.double())
    v.backward()
    o = c_oracle.pattern_loss(disp.detach().numpy(), im_l.numpy(), im_s.numpy(), mod.pattern.numpy(), 9, 3, 0.5, True, "f64")
    assert_scalar_close(o["val"], v.item(), 1e-12)
    assert_close(o["proj"], p, 1e-11)
    assert_close(o["grad_disp"], disp.grad, 1e-10)


def test_smooth_and_warp_c_oracle_f64_equal_reference_f64(ref, frames):
    hw, d, *_ = frames
    rng = np.random.default_rng(0)
    disp = torch.from_numpy(d["disp_gt"] + rng.standard_normal(d["disp_gt"].shape)).double().requires_grad_(True)
    amb = torch.from_numpy(d["ambient"]).double()
    v = ref.networks.DisparitySmoothLoss().double()(disp, amb)
    v.backward()
    ov, og = c_oracle.smooth_loss(disp.detach().numpy(), amb.numpy(), True, "f64")
    assert_scalar_close(ov, v.item(), 1e-12)
    assert_close(og, disp.grad, 1e-11)
    x = torch.randn(3, 5, *hw, dtype=torch.float64, requires_grad=True)
    fl = torch.from_numpy(synth.make_flows(3, hw, max_mag=11.0)[0]).double().requires_grad_(True)
    y = ref.multi_frame_networks.warp(x, fl)
    go = torch.randn_like(y)
    y.backward(go)
    oy, _, _ = c_oracle.flow_warp_forward(x.detach().numpy(), fl.detach().numpy(), "f64")
    gx, gf = c_oracle.flow_warp_backward(x.detach().numpy(), fl.detach().numpy(), go.numpy(), True, "f64")
    assert_close(oy, y, 1e-11)
    assert_close(gx, x.grad, 1e-11)
    assert_close(gf, fl.grad, 1e-10)


def test_lcn_c_oracle_equals_reference(ref):
    x = torch.rand(2, 1, 30, 26)
    for radius in (2, 5):
        mod = ref.networks.LCN(radius, 0.05)
        with torch.no_grad():
            l, s = mod.double()(x.double())
        ol, os_ = c_oracle.lcn_forward(x.numpy(), radius, 0.05, "f64")
        assert_close(ol, l, 1e-13)
        assert_close(os_, s, 1e-13)


@pytest.mark.parametrize("mf", [False, True])
def test_flow_consistency_c_oracle_f64_equals_reference_f64(ref, mf):
    hw = (36, 48)
    g = synth.make_geometry(2, hw, seed=5)
    dt = torch.float64
    K = torch.from_numpy(g["K"].astype(np.float64))
    Ki = torch.from_numpy(np.linalg.inv(g["K"].astype(np.float64)))
    cls = ref.networks.Multi_Frame_Flow_Consistency_Loss if mf else ref.networks.Single_Frame_Flow_Consistency_Loss
    mod = cls(K, Ki, hw[0], hw[1], clamp=0.1)
    mod.ray, mod.u, mod.v = mod.ray.to(dt), mod.u.to(dt), mod.v.to(dt)
    T = lambda k: torch.from_numpy(g[k]).to(dt)
    d0, d1 = T("depth0").requires_grad_(True), T("depth1").requires_grad_(True)
    args = [d0, d1, T("R0"), T("t0"), T("R1"), T("t1"), T("flow01"), T("flow10"), T("amb0"), T("amb1")]
    pd0, pd1 = g["depth0"] + 0.001, g["depth1"] - 0.001
    if mf:
        loss = mod(*args, torch.from_numpy(pd0).to(dt), torch.from_numpy(pd1).to(dt))
    else:
        loss, m0, m1, om = mod(*args)
    loss.backward()
    ray = c_oracle.make_rays(g["K"], *hw)
    kw = dict(clamp=-1.0 if mf else 0.1, prec="f64")
    A = c_oracle.flow_consistency_dir(g["depth0"], g["depth1"], g["R0"], g["t0"], g["R1"], g["t1"], g["flow01"], g["flow10"],
                                      g["amb0"], g["amb1"], g["K"], ray, primary_depth1=pd1 if mf else None, **kw)
    B = c_oracle.flow_consistency_dir(g["depth1"], g["depth0"], g["R1"], g["t1"], g["R0"], g["t0"], g["flow10"], g["flow01"],
                                      g["amb1"], g["amb0"], g["K"], ray, primary_depth1=pd0 if mf else None, **kw)
    assert_scalar_close(A["loss"] + B["loss"], loss.item(), 1e-10)
    assert_close(A["grad_depth0"] + B["grad_depth1"], d0.grad, 1e-9)
    assert_close(A["grad_depth1"] + B["grad_depth0"], d1.grad, 1e-9)
    if not mf:
        assert np.array_equal(A["mask"], m0.numpy()) and np.array_equal(B["mask"], m1.numpy())
        assert np.array_equal(A["orig_mask"][0, 0], om)
    # and the torch port is the same op sequence (fp32, CPU): bit-identical
    f = torch.float32
    modf = cls(K.float(), Ki.float(), hw[0], hw[1], clamp=0.1)
    port = torch_port.FlowConsistency(K.float(), Ki.float(), hw[0], hw[1], clamp=0.1, multi_frame=mf)
    Tf = lambda k: torch.from_numpy(g[k])
    a32 = [Tf("depth0"), Tf("depth1"), Tf("R0"), Tf("t0"), Tf("R1"), Tf("t1"), Tf("flow01"), Tf("flow10"), Tf("amb0"), Tf("amb1")]
    extra = [torch.from_numpy(pd0.astype(np.float32)), torch.from_numpy(pd1.astype(np.float32))] if mf else []
    r, p = modf(*a32, *extra), port(*a32, *extra)
    assert torch.equal(r if mf else r[0], p if mf else p[0])


@pytest.mark.parametrize("stride", [1, 2])
def test_conv3d_gather_port_reproduces_reference_conv3d(ref, stride):
    """The port covers Conv3D.tforward up to the two gathers (model/multi_frame_networks.py:469-501); finishing with
    the module's own MLP / matmul / norm must reproduce the reference layer's output."""
    torch.manual_seed(stride)
    tl, bs, C, h, w = 4, 2, 8, 12, 10
    conv = ref.multi_frame_networks.Conv3D(C, C, stride=stride)
    xyz = torch.randn(tl, bs, 3, h, w) * 0.1
    xyz[:, :, 2] += 1.5
    feat = torch.randn(tl, bs, C, h, w)
    mask = (torch.rand(tl, bs, 1, h, w) > 0.3).float()
    mask[0] = 1.0
    with torch.no_grad():
        out_ref = conv(xyz, feat, mask)
        xyz_nb, feat_nb, ind = torch_port.conv3d_gather(xyz, feat, mask, 3, stride, 9)
        d2 = conv.dense2(conv.dense1(xyz_nb))
        fw = (d2 * feat_nb).sum(dim=1)
        oh, ow = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        out = conv.bn(conv.activation(torch.matmul(fw, conv.w).view(bs, oh, ow, C).permute(0, 3, 1, 2)))
    assert torch.equal(out, out_ref)
