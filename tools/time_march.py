"""Time the 4-scale fused census loss (256 frames of 512x432) under env overrides: python tools/time_march.py [K=V ...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from depthinspace_b200 import _ops, synth  # noqa: E402

for kv in sys.argv[1:]:
    k, v = kv.split("=")
    os.environ[k] = v
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
N, hw = 256, (512, 432)
d = synth.make_frames(8, hw, "default", n_scales=4, max_disp=128, seed=0)
rep = N // 8
lcn_im, std = _ops.lcn_forward(dev(np.tile(d["im"], (rep, 1, 1, 1))), 5, 0.05)
pat = _ops.lcn_forward(dev(d["pattern"]), 5, 0.05)[0].reshape(hw)
disps = [dev(np.tile(p, (rep, 1, 1, 1))) for p in d["disp_pred"]]
best = 1e9
for rnd in range(3):
    for _ in range(3):
        _ops.pattern_loss_multi_forward(disps, lcn_im, std, pat, 9, "census_sad", 0.5, True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10):
        _ops.pattern_loss_multi_forward(disps, lcn_im, std, pat, 9, "census_sad", 0.5, True)
    ev[1].record()
    torch.cuda.synchronize()
    best = min(best, ev[0].elapsed_time(ev[1]) / 10)
print(json.dumps({"args": sys.argv[1:], "lib": os.environ.get("DIS_B200_LIB", "default"), "ms": best}), flush=True)
