import torch, sys, os
sys.path.insert(0, os.getcwd())
from depthinspace_b200 import _ops
x = torch.rand(256, 1, 512, 432, device="cuda")
for _ in range(3): _ops.lcn_forward(x, 5, 0.05)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): _ops.lcn_forward(x, 5, 0.05)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("DIS_B200_LIB","default").split("/")[-1], "lcn ms", e0.elapsed_time(e1)/20)
