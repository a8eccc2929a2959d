"""A/B of two libdis_b200 builds on the LCN outputs: python tools/lcn_ab.py save out.npz | python tools/lcn_ab.py cmp a.npz b.npz"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if sys.argv[1] == "save":
    import torch
    from depthinspace_b200 import _ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(8, 1, 512, 432, device="cuda", generator=g)
    x[1] = 0.37
    x[2, :, :, 200:] += 100.0
    x[3] *= 1e-4
    y = torch.rand(3, 1, 37, 53, device="cuda", generator=g)
    l, s = _ops.lcn_forward(x, 5, 0.05)
    l2, s2 = _ops.lcn_forward(y, 3, 0.05)
    np.savez(sys.argv[2], l=l.cpu().numpy(), s=s.cpu().numpy(), l2=l2.cpu().numpy(), s2=s2.cpu().numpy())
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in a.files:
        d = a[k].view(np.int32).astype(np.int64) - b[k].view(np.int32).astype(np.int64)
        print(k, "bitwise equal" if not d.any() else f"differs: {np.count_nonzero(d)} of {d.size}, max ulp {np.abs(d).max()}")
