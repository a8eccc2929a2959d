"""DIS-MF hot path (BASELINE configs[2]) timed on one GPU: frames/s of everything the path covers in one
multi-frame training step, forward + backward, device-resident inputs, CUDA events.

Per track of tl = 4 frames (reference call sites in brackets):
  * copy_data:   transpose + LCN + cat of the IR frames                       [worker.py:418-452]
  * FuseNet:     24 xyz / flow warps with forward-backward masks at 256x216   [multi_frame_networks.py:187-214]
                 96 feature warps, C = 32: 4 blocks x 4 frames x 3 neighbours at 256x216 and 128x108 [:347-360]
                 each run forward, forward again (torch.utils.checkpoint recompute) and backward w.r.t. x
  * loss:        1-scale census_sad 9x9 pattern loss + 0.8 smoothness + 6 pairs x 2 directions of the multi-frame
                 flow-consistency loss + primary-disparity L1, forward + backward   [multi_frame_worker.py:103-175]
The FuseNet convolutions / Conv3D MLPs stay on PyTorch (north star) and are not part of the timed region.

    python tools/bench_mf.py [--bs 4 32] [--steps 10] [--warmup 3]

One JSON line per batch size; `frames` = tl * bs.  bs = 4 is one GPU's share of configs[2] (batch 32 over 8 GPUs).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depthinspace_b200 import _lib, losses, multi_frame_networks as mfn, networks, synth  # noqa: E402

TL = 4
HW = synth.DATASET_HW
FEAT = ((HW[0] // 2, HW[1] // 2), (HW[0] // 4, HW[1] // 4))   # 256x216, 128x108
C_FEAT, BLOCKS = 32, 4
ALL_FRAMES = True
CONV3D = False
RANK_ONCE = True


def build(bs, dev, group=None):
    n = TL * bs
    base = min(n, 8)                      # a few distinct synthetic frames, tiled up to the batch
    fr = synth.make_frames(base, HW, "default", n_scales=1, seed=42)
    rep = lambda a: torch.from_numpy(np.concatenate([a] * ((n + base - 1) // base))[:n]).to(dev)
    im, amb, disp, dgt = rep(fr["im"]), rep(fr["ambient"]), rep(fr["disp_pred"][0]), rep(fr["disp_gt"])
    g = synth.make_geometry(min(bs, 2), HW, seed=5)
    repb = lambda a: torch.from_numpy(np.concatenate([a] * bs)[:bs]).to(dev)
    K = torch.from_numpy(g["K"].astype(np.float64))
    Ki = torch.from_numpy(np.linalg.inv(g["K"].astype(np.float64)))
    # every ordered frame pair of the track uses the two-view geometry (even frames = view 0, odd = view 1)
    R = torch.stack([repb(g["R0"] if i % 2 == 0 else g["R1"]) for i in range(TL)])
    t = torch.stack([repb(g["t0"] if i % 2 == 0 else g["t1"]) for i in range(TL)])
    flow = {}
    for i in range(TL):
        for j in range(TL):
            if i != j:
                same = (i % 2) == (j % 2)
                f = np.zeros_like(g["flow01"]) if same else (g["flow01"] if i % 2 == 0 else g["flow10"])
                flow[f"flow_{i}{j}"] = repb(f)
    view = lambda a: a.view(TL, bs, *a.shape[1:])
    pattern = torch.from_numpy(np.repeat(fr["pattern"], 3, axis=1)).to(dev)
    focal, baseline = float(g["K"][0, 0]), 0.075
    loss = losses.MultiFrameLoss(HW[0], HW[1], pattern, K=K, Ki=Ki, focal_length=focal, baseline=baseline,
                                 process_group=group).to(dev)
    lcn = networks.LCN(5, 0.05).to(dev)
    feats, grads, grads_all = [], [], []
    gen = torch.Generator(device=dev).manual_seed(0)
    for (h, w) in FEAT:
        feats.append(torch.randn(TL, bs, C_FEAT, h, w, device=dev, generator=gen).requires_grad_(True))
        grads.append(torch.randn(TL, bs, C_FEAT, h, w, device=dev, generator=gen))
        grads_all.append(grads[-1][None].expand(TL, -1, -1, -1, -1, -1).contiguous())
    xyz = torch.randn(TL, bs, 3, *FEAT[0], device=dev, generator=gen)
    xyz_lvl, mask_lvl, g_nb = [], [], []
    if CONV3D:
        for (h, w) in FEAT:
            v, u = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32), torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
            xl = torch.empty(TL, bs, 3, h, w, device=dev)
            for fr_i in range(TL):
                z = 1.5 + 0.2 * torch.sin(u / 40 + fr_i) * torch.cos(v / 55) + 0.002 * torch.randn(bs, h, w, device=dev, generator=gen)
                xl[fr_i, :, 0], xl[fr_i, :, 1], xl[fr_i, :, 2] = (u - w / 2 + 0.3 * fr_i) / (1.3 * w) * z, (v - h / 2 - 0.2 * fr_i) / (1.3 * w) * z, z
            xyz_lvl.append(xl)
            mask_lvl.append((torch.rand(TL, bs, 1, h, w, device=dev, generator=gen) > 0.1).float())
            g_nb.append(torch.randn(bs * h * w, 9, C_FEAT, device=dev, generator=gen))
    im_bt = view(im).transpose(0, 1).contiguous()      # the DataLoader hands frames over as [bs, tl, 1, H, W]
    return dict(im=view(im), im_bt=im_bt, amb=view(amb), disp=view(disp), prim=view(dgt + 0.3), R=R, t=t, flow=flow, loss=loss, lcn=lcn,
                feats=feats, grads=grads, grads_all=grads_all, xyz=xyz, bs=bs, xyz_lvl=xyz_lvl, mask_lvl=mask_lvl, g_nb=g_nb)


def resize_flows(w):
    """FuseNet's resize_flow_like (multi_frame_networks.py:58-72) for the two feature resolutions: every step gets the
    flows at full resolution from the data loader and needs them at 256x216 and 128x108."""
    return [mfn.resize_flow_like(w["flow"], size) for size in FEAT]


class Sections:
    """CUDA-event stopwatch per section of the step (accumulated over the timed steps)."""

    def __init__(self):
        self.marks, self.on = [], False

    def mark(self, name):
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def summary(self, steps):
        acc = {}
        for (_, e0), (name, e1) in zip(self.marks, self.marks[1:]):
            if name != "start":
                acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e1)
        return {k: round(v / steps, 3) for k, v in acc.items()}


SEC = Sections()


def step(w):
    SEC.mark("start")
    # ---- copy_data: LCN + cat(lcn, raw), [tl, bs, 2, H, W]
    im_cat, std = w["lcn"].prepare_input(w["im_bt"])
    SEC.mark("copy_data_ms")
    # ---- FuseNet warps
    w["flows_lr"] = resize_flows(w)
    SEC.mark("resize_flows_ms")
    with torch.no_grad():
        for tidx in range(TL):
            mfn.gather_warped(w["xyz"], w["flows_lr"][0], tidx, with_fb_mask=True)          # 12 xyz + 12 flow warps
    SEC.mark("xyz_flow_warps_ms")
    for lvl in range(2):
        x, fl, go = w["feats"][lvl], w["flows_lr"][lvl], w["grads"][lvl]
        x.grad = None
        for _ in range(BLOCKS):
            if ALL_FRAMES:       # one gather for all target frames (the tidx loop of fwd_3d_1 / fwd_3d_2 as one op)
                if CONV3D and RANK_ONCE and _ == 0:
                    # the selection depends on the level's xyz / mask only: ranked once per (level, target frame) and step,
                    # shared by the Conv3D layers of the level and by the checkpoint recompute
                    ranks = [mfn.conv3d_rank(w["xyz_lvl"][lvl], w["mask_lvl"][lvl], 3, 1, 9) for tidx in range(TL)]
                rk = (lambda t: ranks[t]) if (CONV3D and RANK_ONCE) else (lambda t: None)
                with torch.no_grad():
                    warped = mfn.gather_warped_all(x, fl)
                    if CONV3D:   # Conv3D's neighbour selection + gather on every target frame (checkpointed forward)
                        for tidx in range(TL):
                            mfn.conv3d_gather(w["xyz_lvl"][lvl], warped[tidx], w["mask_lvl"][lvl], 3, 1, 9, rank=rk(tidx))
                warped = mfn.gather_warped_all(x, fl)
                if CONV3D:
                    g_acc = torch.zeros_like(warped)
                    for tidx in range(TL):
                        wf = warped[tidx].detach().requires_grad_(True)
                        _, feat_nb, _ = mfn.conv3d_gather(w["xyz_lvl"][lvl], wf, w["mask_lvl"][lvl], 3, 1, 9, rank=rk(tidx))
                        feat_nb.backward(w["g_nb"][lvl])
                        g_acc[tidx] = wf.grad
                    warped.backward(g_acc)
                else:
                    warped.backward(w["grads_all"][lvl])
                continue
            for tidx in range(TL):
                with torch.no_grad():
                    mfn.gather_warped(x, fl, tidx)                                          # checkpointed forward
                out = mfn.gather_warped(x, fl, tidx)                                        # recompute in backward
                out.backward(go)
        SEC.mark(f"feature_warps_{FEAT[lvl][0]}x{FEAT[lvl][1]}_ms")
    # ---- loss
    disp = w["disp"].detach().requires_grad_(True)
    vals = w["loss"]([disp], im_cat, std, w["amb"], primary_disp=w["prim"], R=w["R"], t=w["t"], flow_out=w["flow"])
    total = torch.stack([v.reshape(()) for v in vals]).sum()
    total.backward()
    SEC.mark("loss_ms")
    return total


def build_sf(bs, dev):
    """DIS-SF loss with its geometric terms (single_frame_worker.py:101-149): 4 scales + smoothness + 6 pairs x 2 directions."""
    w = build(bs, dev)
    n = TL * bs
    fr = synth.make_frames(min(n, 8), HW, "default", n_scales=4, seed=42)
    rep = lambda a: torch.from_numpy(np.concatenate([a] * ((n + 7) // 8))[:n]).to(dev)
    w["disps"] = [rep(p).view(TL, bs, 1, *HW) for p in fr["disp_pred"]]
    old = w["loss"]
    pattern = old.ph_loss.pattern.repeat(1, 3, 1, 1)
    g = synth.make_geometry(1, HW, seed=5)
    K = torch.from_numpy(g["K"].astype(np.float64))
    w["loss"] = losses.SingleFrameLoss(HW[0], HW[1], pattern, K=K, Ki=torch.linalg.inv(K), focal_length=float(g["K"][0, 0]),
                                       baseline=0.075).to(dev)
    for k in ("feats", "grads", "grads_all", "xyz", "xyz_lvl", "mask_lvl", "g_nb"):
        w.pop(k)
    return w


def step_sf(w):
    SEC.mark("start")
    im_cat, std = w["lcn"].prepare_input(w["im_bt"])
    SEC.mark("copy_data_ms")
    disps = [d.detach().requires_grad_(True) for d in w["disps"]]
    vals = w["loss"](disps, im_cat, std, w["amb"], R=w["R"], t=w["t"], flow_out=w["flow"])
    total = torch.stack([v.reshape(()) for v in vals]).sum()
    total.backward()
    SEC.mark("loss_ms")
    return total


WARP_BYTES_PER_FRAME = 657e6   # warp traffic of one FuseNet fwd + recompute + bwd per frame


def cpu_step(threads):
    """The same hot path for ONE track (bs 1, 4 frames) with the oracle's torch port on the host cores (the reference's
    op sequence: LCN, warp() per neighbour + torch.stack, unfold-based photometric loss, flow-consistency modules)."""
    import time
    from oracle import torch_port as tp
    torch.set_num_threads(threads)
    w = build(1, torch.device("cpu"))
    t0 = time.perf_counter()
    im = w["im"].reshape(TL, 1, *HW)
    im_l, im_s = tp.lcn(im)
    fl_lr = [{k: torch.nn.functional.interpolate(v, size=size, mode="bilinear", align_corners=True) * (size[0] / HW[0])
              for k, v in w["flow"].items()} for size in FEAT]
    with torch.no_grad():
        for t in range(TL):
            for j in range(TL):
                if j != t:
                    tp.flow_warp(w["xyz"][j], fl_lr[0][f"flow_{t}{j}"])
                    tp.fb_mask(fl_lr[0][f"flow_{t}{j}"], tp.flow_warp(fl_lr[0][f"flow_{j}{t}"], fl_lr[0][f"flow_{t}{j}"]))
    for lvl in range(2):
        x = w["feats"][lvl]
        x.grad = None
        for _ in range(BLOCKS):
            for t in range(TL):
                def gather():
                    return torch.stack([x[t]] + [tp.flow_warp(x[j], fl_lr[lvl][f"flow_{t}{j}"]) for j in range(TL) if j != t])
                with torch.no_grad():
                    gather()                                   # checkpointed forward
                gather().backward(w["grads"][lvl])            # recompute + backward
    disp = w["disp"].detach().reshape(TL, 1, *HW).requires_grad_(True)
    pat = w["loss"].ph_loss.pattern
    vals = tp.multi_frame_loss(disp, im_l, im_s, w["amb"].reshape(TL, 1, *HW), pat, chunk=1, primary_disp=w["prim"].reshape(TL, 1, *HW))
    g = synth.make_geometry(1, HW, seed=5)
    K = torch.from_numpy(g["K"])
    fc = tp.FlowConsistency(K, torch.linalg.inv(K.double()).float(), HW[0], HW[1], multi_frame=True)
    depth = tp.disp_to_depth(disp.view(TL, 1, 1, *HW), float(g["K"][0, 0]), 0.075)
    prim = tp.disp_to_depth(w["prim"], float(g["K"][0, 0]), 0.075)
    for i in range(TL):
        for j in range(i + 1, TL):
            vals.append(fc(depth[i], depth[j], w["R"][i], w["t"][i], w["R"][j], w["t"][j], w["flow"][f"flow_{i}{j}"], w["flow"][f"flow_{j}{i}"],
                           w["amb"][i], w["amb"][j], prim[i], prim[j]) * 0.2 / 6)
    sum(vals).backward()
    return time.perf_counter() - t0


def measure(bs, dev, group, world, steps, peak_gbs, sync_all, cpu_leg):
    """DIS-MF numbers for bench.py: `bs` samples on this rank (BASELINE configs[2]: 32 in total), frames/s over all ranks
    (max over ranks of the device time), the dominant kernel against the HBM roofline, the end-to-end leg and (one GPU
    only) the CPU leg."""
    import time
    import torch.distributed as dist
    from depthinspace_b200 import _ops
    w = build(bs, dev, group)
    for _ in range(3):
        step(w)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    SEC.marks, SEC.on = [], True
    e0.record()
    for _ in range(steps):
        total = step(w)
    e1.record()
    sync_all()
    SEC.on = False
    ms = e0.elapsed_time(e1)
    sections = SEC.summary(steps)
    # dominant kernel alone: backward of the all-frames feature gather at 256x216, C = 32
    h, wd = FEAT[0]
    g = w["grads_all"][0]
    fl = {(i, j): w["flows_lr"][0][f"flow_{i}{j}"] for i in range(TL) for j in range(TL) if i != j}
    for _ in range(3):
        _ops.flow_warp_gather_all_backward(fl, g)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(10):
        _ops.flow_warp_gather_all_backward(fl, g)
    k1.record()
    torch.cuda.synchronize()
    k_ms = k0.elapsed_time(k1) / 10
    k_bytes = (TL * TL * C_FEAT + TL * (TL - 1) * 2 + TL * C_FEAT) * 4 * h * wd * bs
    # end to end: the step's data-loader inputs come from pinned host memory every step, the loss goes back
    keys = ("im_bt", "amb", "disp", "prim", "R", "t")
    host = {k: w[k].cpu().pin_memory() for k in keys}
    host_flow = {k: v.cpu().pin_memory() for k, v in w["flow"].items()}
    h2d = sum(v.numel() * 4 for v in host.values()) + sum(v.numel() * 4 for v in host_flow.values())
    side = torch.cuda.Stream()

    def upload():
        with torch.cuda.stream(side):
            bufs = ({k: v.to(dev, non_blocking=True) for k, v in host.items()}, {k: v.to(dev, non_blocking=True) for k, v in host_flow.items()})
            ev = torch.cuda.Event()
            ev.record(side)
        return bufs, ev

    def run(k):
        nxt = upload()
        out = []
        for i in range(k):
            (b, f), ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            if i + 1 < k:
                nxt = upload()
            for t in list(b.values()) + list(f.values()):
                t.record_stream(torch.cuda.current_stream())
            w.update(b)
            w["flow"] = f
            out.append(float(step(w).detach()))
        return out
    run(2)
    sync_all()
    t0 = time.perf_counter()
    run(steps)
    sync_all()
    e2e_s = time.perf_counter() - t0
    times = torch.tensor([ms, e2e_s * 1e3, k_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, k_ms = times.tolist()
    frames = TL * bs * world
    value = frames * steps / (ms * 1e-3)
    line = {"value": value, "unit": "frames/s", "ms_per_step": ms / steps, "frames_per_step": frames, "samples_per_gpu": bs, "scaling": "strong",
            "e2e": {"value": frames * steps / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "sections_ms": sections,
            "roofline": {"bound": "hbm", "kernel": "flow_warp_gather_all_bwd_kernel (tl 4, C 32, 256x216)", "kernel_ms": k_ms,
                         "algorithmic_bytes": k_bytes, "achieved": k_bytes / (k_ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                         "frac": k_bytes / (k_ms * 1e-3) / 1e9 / peak_gbs,
                         "step_algorithmic_gbs": WARP_BYTES_PER_FRAME * (value / world) / 1e9,
                         "step_frac": WARP_BYTES_PER_FRAME * (value / world) / 1e9 / peak_gbs,
                         "limiter": "L2 reduction path of the feature-gradient scatter (RED.ADD), see DESIGN.md"},
            "workload": "BASELINE configs[2]: DIS-MF hot path, bs 32 in total x tl 4 (copy_data LCN, flow resize to 2 levels, 24 xyz/flow + 96 C=32 "
                        "feature warps fwd/recompute/bwd, 1-scale census_sad + smoothness + 12 flow-consistency terms + L1), see tools/bench_mf.py",
            "loss": float(total.detach())}
    del w
    torch.cuda.empty_cache()
    if cpu_leg:
        threads = os.cpu_count() or 1
        t = cpu_step(threads)
        line["cpu_baseline"] = {"value": TL / t, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": "1 track (4 frames) of the 32, 1 pass, oracle torch port, all host threads"}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, nargs="+", default=[4, 32])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--per-frame-gathers", action="store_true", help="one gather call per target frame (reference loop shape)")
    ap.add_argument("--conv3d-rank-per-call", action="store_true", help="with --conv3d: rank the neighbours in every call (round-1 behaviour) "
                    "instead of once per level, target frame and step")
    ap.add_argument("--conv3d", action="store_true", help="also run Conv3D's neighbour selection + gather (SURVEY 8(f) row 3) on every "
                    "gathered stack: forward, checkpoint recompute and backward")
    ap.add_argument("--sf", action="store_true", help="time the DIS-SF loss WITH its 12 flow-consistency terms instead (bs 64 = 256 frames)")
    a = ap.parse_args()
    global ALL_FRAMES, CONV3D, RANK_ONCE
    ALL_FRAMES = not a.per_frame_gathers
    CONV3D = a.conv3d
    RANK_ONCE = not a.conv3d_rank_per_call
    dev = torch.device("cuda")
    run_step = step_sf if a.sf else step
    if a.sf and a.bs == [4, 32]:
        a.bs = [64]
    for bs in a.bs:
        w = build_sf(bs, dev) if a.sf else build(bs, dev)
        for _ in range(a.warmup):
            total = run_step(w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        l0 = _lib.LAUNCHES
        SEC.marks, SEC.on = [], True
        e0.record()
        for _ in range(a.steps):
            total = run_step(w)
        e1.record()
        torch.cuda.synchronize()
        SEC.on = False
        ms = e0.elapsed_time(e1) / a.steps
        launches = (_lib.LAUNCHES - l0) // a.steps
        n = TL * bs
        name = ("DIS-SF loss path with geometric terms (copy_data LCN + 4 x census_sad 9x9 + smoothness + 12 flow-consistency terms), fwd+bwd"
                if a.sf else
                "BASELINE configs[2]: DIS-MF hot path (copy_data LCN + 24 xyz/flow warps + 96 C=32 feature warps "
                "fwd/recompute/bwd + 1-scale census_sad loss + smoothness + 12 flow-consistency terms + L1), fwd+bwd")
        print(json.dumps({"workload": name,
                          "gather": "all frames per call" if ALL_FRAMES else "one call per target frame", "conv3d_gather": CONV3D, "bs": bs, "tl": TL, "frames": n, "ms_per_step": round(ms, 3), "frames_per_s": round(n / (ms * 1e-3), 1),
                          "gpu_launches_per_step": launches, "sections": SEC.summary(a.steps), "loss": float(total.detach())}), flush=True)
        del w
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
