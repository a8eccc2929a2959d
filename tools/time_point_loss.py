"""Time the mse / sad point-wise pattern loss (256 frames of 512x432): box passes vs point kernel, S = 1..4."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depthinspace_b200 import _ops
N, H, W = 256, 512, 432
g = torch.Generator(device="cuda").manual_seed(0)
im = torch.randn(N, 1, H, W, device="cuda", generator=g)
std = torch.rand(N, 1, H, W, device="cuda", generator=g) + 0.05
pat = torch.randn(H, W, device="cuda", generator=g)
d4 = [torch.rand(N, 1, H, W, device="cuda", generator=g) * 60 for _ in range(4)]
ws = torch.empty(2 * im.numel(), device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for lt in ("mse", "sad"):
    for S in (1, 2, 4):
        full = t(lambda: _ops.pattern_loss_point_forward(d4[:S], im, std, pat, 9, lt, False, True, workspace=ws))
        reuse = t(lambda: _ops.pattern_loss_point_forward(d4[:S], im, std, pat, 9, lt, False, True, workspace=ws, reuse_wbox=True))
        print(f"{lt} S={S}: box+point {full:.3f} ms, point only {reuse:.3f} ms, box passes {full - reuse:.3f} ms", flush=True)
