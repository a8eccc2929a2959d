"""Runs every DIS-MF kernel of the library twice on BASELINE-shaped inputs (target for `ncu --set full`):
single flow warp, FuseNet gathers (one frame / all frames), flow-consistency loss, Conv3D neighbour gather,
single-scale census pattern loss.  python tools/run_mf_kernels.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depthinspace_b200 import _ops, networks, synth

dev = torch.device("cuda")
tl, bs, C, h, w = 4, 32, 32, 256, 216
H, W = synth.DATASET_HW
feat = torch.randn(tl, bs, C, h, w, device=dev)
fl = {(i, j): torch.from_numpy(synth.make_flows(bs, (h, w), max_mag=6.0, seed=3 * i + j)[0]).to(dev)
      for i in range(tl) for j in range(tl) if i != j}
go6 = torch.randn(tl, tl, bs, C, h, w, device=dev)
gm = synth.make_geometry(2, (H, W), seed=1)
rep = lambda k: torch.from_numpy(np.concatenate([gm[k]] * 32)).to(dev)
K = torch.from_numpy(gm["K"].astype(np.float64))
fc = networks.Single_Frame_Flow_Consistency_Loss(K, torch.linalg.inv(K), H, W, clamp=0.1)
fc_args = [rep(k) for k in ("depth0", "depth1", "R0", "t0", "R1", "t1", "flow01", "flow10", "amb0", "amb1")]
fc_args[0].requires_grad_(True); fc_args[1].requires_grad_(True)
v, u = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32), torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
xyz = torch.empty(tl, 4, 3, h, w, device=dev)
for t in range(tl):
    z = 1.5 + 0.2 * torch.sin(u / 40 + t) * torch.cos(v / 55) + 0.002 * torch.randn(4, h, w, device=dev)
    xyz[t, :, 0], xyz[t, :, 1], xyz[t, :, 2] = (u - w / 2 + 0.3 * t) / 280 * z, (v - h / 2 - 0.2 * t) / 280 * z, z
mask = (torch.rand(tl, 4, 1, h, w, device=dev) > 0.1).float()
fr = synth.make_frames(8, (H, W), "default", n_scales=1, seed=4)
im = torch.from_numpy(np.concatenate([fr["im"]] * 16)).to(dev)
disp = torch.from_numpy(np.concatenate([fr["disp_pred"][0]] * 16)).to(dev)
lcn = networks.LCN(5, 0.05)
im_l, im_s = lcn(im)
pat_l, _ = lcn(torch.from_numpy(fr["pattern"]).to(dev))
for _ in range(2):
    _ops.flow_warp_forward(feat[0], fl[(0, 1)])
    _ops.flow_warp_backward(None, fl[(0, 1)], go6[0, 0], True, False)
    _ops.flow_warp_gather_all_forward(feat, fl)
    _ops.flow_warp_gather_all_backward(fl, go6)
    fc(*fc_args)[0].backward()
    xyz_nb, feat_nb, idx, _ = _ops.conv3d_gather_forward(xyz, feat[:, :4].contiguous(), mask, 3, 1, 9)
    _ops.conv3d_gather_backward(None, torch.ones_like(feat_nb), idx, (tl, 4, C, h, w), 3, 1, 9, False, True)
    _ops.pattern_loss_forward(disp, im_l, im_s, pat_l, 9, "census_sad", 0.5, False, False, True)
torch.cuda.synchronize()
print("done")
