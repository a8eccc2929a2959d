"""BASELINE configs[4]: window / LCN-radius / disparity-range sweep of the DIS-SF loss path on the real dot pattern
(one GPU, 256 frames of 512x432, device-resident, CUDA events).  One JSON line per configuration.
    python tools/bench_sweep.py [--frames 256] [--steps 5]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depthinspace_b200 import losses, networks, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda")
    hw = synth.DATASET_HW
    n = a.frames
    for max_disp in (64.0, 128.0, 192.0):
        fr = synth.make_frames(8, hw, "real", n_scales=4, max_disp=max_disp, seed=42)
        rep = lambda x: torch.from_numpy(np.concatenate([x] * (n // 8))).to(dev)
        im, amb = rep(fr["im"]), rep(fr["ambient"])
        disps = [rep(p) for p in fr["disp_pred"]]
        pat = torch.from_numpy(fr["pattern"]).to(dev)
        for radius in (3, 5, 7):
            lcn = networks.LCN(radius, 0.05)
            pat_l, _ = lcn(pat)
            for k in (5, 7, 9, 11, 13):
                if max_disp != 128.0 and (radius != 5 or k != 9):
                    continue      # the disparity range only changes the data, not the work: one point each
                loss = losses.SingleFrameLoss(hw[0], hw[1], torch.cat([pat_l] * 3, dim=1), block_size=k)

                def step():
                    im_l, im_s = lcn(im)
                    vals, grads = loss.value_and_grad(disps, im_l, im_s, amb)
                    return torch.stack(vals).sum()
                for _ in range(2):
                    total = step()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.steps):
                    total = step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.steps
                print(json.dumps({"pattern": "real", "max_disp": max_disp, "lcn_radius": radius, "block_size": k, "frames": n,
                                  "ms_per_step": round(ms, 3), "frames_per_s": round(n / (ms * 1e-3), 1),
                                  "loss": round(float(total), 6)}), flush=True)


if __name__ == "__main__":
    main()
