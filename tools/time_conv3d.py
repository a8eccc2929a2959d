"""Runs the Conv3D neighbour gather (forward + feature backward) three times on realistic geometry; meant to be run under
`ncu --metrics gpu__time_duration.sum -k regex:conv3d` for per-kernel times."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depthinspace_b200 import _ops
tl, bs, C, h, w = 4, 4, 32, 256, 216
dev = "cuda"
# realistic geometry: points on a smooth surface seen by 4 slightly displaced frames
v, u = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32), torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
f = 1.3 * w
xyz = torch.empty(tl, bs, 3, h, w, device=dev)
for t in range(tl):
    z = 1.5 + 0.2 * torch.sin(u / 40 + t) * torch.cos(v / 55) + 0.002 * torch.randn(bs, h, w, device=dev)
    xyz[t, :, 0] = (u - w / 2 + 0.3 * t) / f * z
    xyz[t, :, 1] = (v - h / 2 - 0.2 * t) / f * z
    xyz[t, :, 2] = z
feat = torch.randn(tl, bs, C, h, w, device=dev)
mask = (torch.rand(tl, bs, 1, h, w, device=dev) > 0.1).float()
for _ in range(3):
    xyz_nb, feat_nb, idx, _ = _ops.conv3d_gather_forward(xyz, feat, mask, 3, 1, 9)
    g = torch.randn_like(feat_nb)
    _ops.conv3d_gather_backward(None, g, idx, (tl, bs, C, h, w), 3, 1, 9, False, True)
torch.cuda.synchronize()
print("distinct candidate ids per slot (sample):", [int(idx[:, j].unique().numel()) for j in range(9)])
