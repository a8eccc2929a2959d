"""Which of the ops can be captured into a CUDA graph?  (diagnostic; run on the GPU box)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depthinspace_b200 import _ops, networks, synth

dev = torch.device("cuda")
hw = (128, 160)
d = synth.make_frames(4, hw, n_scales=4)
T = lambda a: torch.from_numpy(a).to(dev)
im, amb, pat = T(d["im"]), T(d["ambient"]), T(d["pattern"])
disps = [T(p) for p in d["disp_pred"]]
iml, ims = _ops.lcn_forward(im, 5, 0.05)
one = torch.ones(1, device=dev)

def probe(name, fn):
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name)
    except Exception as e:
        print("FAIL", name, str(e).split("\n")[0][:120])
        torch.cuda.synchronize()

probe("lcn", lambda: _ops.lcn_forward(im, 5, 0.05))
probe("smooth", lambda: _ops.smooth_loss_forward(disps[0], amb, True))
probe("pattern single", lambda: _ops.pattern_loss_forward(disps[0], iml, ims, pat, 9, 3, 0.5, False, False, True))
probe("pattern multi", lambda: _ops.pattern_loss_multi_forward(disps, iml, ims, pat, 9, 3, 0.5, True))
probe("scale", lambda: _ops.scale_by_device_scalar(disps[0], one, one))
probe("photometric fwd", lambda: _ops.photometric_loss_forward(iml, iml, 9, 3, 0.5))

from depthinspace_b200 import losses
lcn = networks.LCN(5, 0.05)
loss = losses.SingleFrameLoss(hw[0], hw[1], torch.cat([_ops.lcn_forward(pat, 5, 0.05)[0]] * 3, 1))
dd = [p.clone().requires_grad_(True) for p in disps]

def fwd_only():
    a, b = lcn(im)
    return sum(loss(dd, a, b, amb))

def fwd_bwd():
    for p in dd:
        p.grad = None
    fwd_only().backward()

def single_bwd():
    dd[0].grad = None
    v, _ = loss.ph_loss(dd[0], iml, ims)
    v.backward()

def smooth_bwd():
    dd[0].grad = None
    loss.disparity_loss(dd[0], amb).backward()

probe("assembly fwd", fwd_only)
probe("single-scale fwd+bwd", single_bwd)
probe("smooth fwd+bwd", smooth_bwd)
probe("assembly fwd+bwd", fwd_bwd)
