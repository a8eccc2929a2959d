"""One all-frames gather forward at the DIS-MF shape (tl 4, bs 32, C 32, 256x216) for ncu: python tools/run_gather_fwd_once.py [bs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_mf  # noqa: E402
from depthinspace_b200 import _ops  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
w = bench_mf.build(bs, torch.device("cuda"))
w["flows_lr"] = bench_mf.resize_flows(w)
fl = {(i, j): w["flows_lr"][0][f"flow_{i}{j}"] for i in range(4) for j in range(4) if i != j}
x = w["feats"][0].detach()
for _ in range(3):
    _ops.flow_warp_gather_all_forward(x, fl)
torch.cuda.synchronize()
