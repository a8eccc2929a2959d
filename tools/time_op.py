"""Time one op on the GPU: python tools/time_op.py smooth|lcn|multi"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depthinspace_b200 import _ops
N, H, W = 256, 512, 432
x = torch.rand(N, 1, H, W, device="cuda")
d = torch.rand(N, 1, H, W, device="cuda") * 60
ops = {"smooth": lambda: _ops.smooth_loss_forward(d, x, True), "smooth_fwd": lambda: _ops.smooth_loss_forward(d, x, False),
       "lcn": lambda: _ops.lcn_forward(x, 5, 0.05)}
if any(a.startswith("gather") for a in sys.argv[1:]):
    tl, bs, C, h, w = 4, 32, 32, 256, 216
    xf = torch.randn(tl, bs, C, h, w, device="cuda")
    go = torch.randn(tl, bs, C, h, w, device="cuda")
    yy, xx = torch.meshgrid(torch.linspace(0, 6.28, h, device="cuda"), torch.linspace(0, 6.28, w, device="cuda"), indexing="ij")
    flows = [torch.stack((4 * torch.sin(xx + k) + 2 * torch.cos(yy), 3 * torch.cos(yy * 2 + k) - torch.sin(xx)), 0)[None].repeat(bs, 1, 1, 1).contiguous()
             for k in range(tl - 1)]
    ops["gather_fwd"] = lambda: _ops.flow_warp_gather_forward(xf, flows, 1)
    ops["gather_bwd"] = lambda: _ops.flow_warp_gather_backward(flows, go, 1)
    ops["gather_warp_bwd"] = lambda: _ops.flow_warp_backward(None, flows[0], go[0], True, False)
for name in sys.argv[1:]:
    fn = ops[name]
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "ms", round(e0.elapsed_time(e1) / 20, 4))
