"""One launch sequence of the 4-scale fused census loss at the bench shape (for ncu): python tools/run_march_once.py [frames]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from depthinspace_b200 import _ops, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = (512, 432)
d = synth.make_frames(8, hw, "default", n_scales=4, max_disp=128, seed=0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rep = N // 8
lcn_im, std = _ops.lcn_forward(dev(np.tile(d["im"], (rep, 1, 1, 1))), 5, 0.05)
pat = _ops.lcn_forward(dev(d["pattern"]), 5, 0.05)[0].reshape(hw)
disps = [dev(np.tile(p, (rep, 1, 1, 1))) for p in d["disp_pred"]]
for _ in range(3):
    _ops.pattern_loss_multi_forward(disps, lcn_im, std, pat, 9, "census_sad", 0.5, True)
torch.cuda.synchronize()
