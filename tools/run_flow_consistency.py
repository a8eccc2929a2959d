"""Run the flow-consistency loss a few times (target for ncu): python tools/run_flow_consistency.py [bs]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depthinspace_b200 import networks, synth
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, W = synth.DATASET_HW
gm = synth.make_geometry(2, (H, W), seed=1)
dev = torch.device("cuda")
rep = lambda k: torch.from_numpy(np.concatenate([gm[k]] * (bs // 2))).to(dev)
K = torch.from_numpy(gm["K"].astype(np.float64))
mod = networks.Single_Frame_Flow_Consistency_Loss(K, torch.linalg.inv(K), H, W, clamp=0.1)
args = [rep(k) for k in ("depth0", "depth1", "R0", "t0", "R1", "t1", "flow01", "flow10", "amb0", "amb1")]
args[0].requires_grad_(True); args[1].requires_grad_(True)
for _ in range(4):
    loss = mod(*args)[0]
    loss.backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    loss = mod(*args)[0]; loss.backward()
e1.record(); torch.cuda.synchronize()
print("flow consistency pair fwd+bwd ms", e0.elapsed_time(e1) / 10, "bs", bs, "loss", float(loss.detach()))
