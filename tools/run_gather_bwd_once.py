"""One all-frames gather backward at the DIS-MF shape (tl 4, bs 32, C 32, 256x216) for ncu: python tools/run_gather_bwd_once.py [bs]"""
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
g = w["grads_all"][0]
for _ in range(3):
    _ops.flow_warp_gather_all_backward(fl, g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _ops.flow_warp_gather_all_backward(fl, g)
e1.record()
torch.cuda.synchronize()
print("gather_all backward ms", e0.elapsed_time(e1) / 10, "impl", os.environ.get("DIS_GATHER_BWD", "tile"))
