#!/usr/bin/env python
"""bench.py -- frames/s of the DepthInSpace self-supervision hot path (loss + warp, fwd + bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic structured-light frames:
    LCN(IR)  ->  4 x RectifiedPatternSimilarityLoss (census_sad 9x9, sigma-weighted, scales 1/2^s)
             ->  DisparitySmoothLoss(scale 0) * 0.4  ->  backward to the 4 disparity maps
i.e. BASELINE.json configs[1] "DIS-SF training step ... batch 64": 64 samples x 4 frames = 256 frames of
512x432 per GPU (the DispNet convolutions are out of scope and stay on cuDNN; they are not in the step).
Frames shard by sample with no data-path collective; with N > 1 ranks the only exchange is the NCCL
all-reduce of the loss numerators / denominators.  "scaling": "weak" (256 frames per GPU).

Printed JSON (rank 0, one line): the driver contract plus
    roofline      dominant kernel (fused pattern-loss) algorithmic bytes / CUDA-event time vs measured HBM peak
    cpu_baseline  the oracle's torch port of the reference timed on this box's host cores (bounded sample)
    e2e           same step through the public modules with HOST (pinned) inputs, H2D + D2H inside the timed region
--impl reference times the reference's CPU implementation of the path (oracle port; the reference is pure
Python and its native dependency is un-vendored, so there is no oracle/_ref binary) on rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/s loss+warp fwd+bwd (DIS-SF)"
UNIT = "frames/s"
HW = (512, 432)
N_SCALES = 4
FRAMES_PER_GPU = 256          # bs 64 x track length 4
ALGO_BYTES_PER_FRAME = 144    # x P, SURVEY.md section 8(d): 12P LCN + 4 x 28P photometric + 20P smoothness
KERNEL_ALGO_BYTES_PER_FRAME = 40  # x P, 4-scale fused pattern-loss kernel: reads 4 disp + im + sigma (24P), writes 4 x d/d disp (16P)


def measured_traffic(kernel, frames):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return int(json.load(f)[kernel]["bytes_per_frame"] * frames)
    except Exception:
        return None


def measured_pipes(kernel):
    """Pipe utilisation of the dominant kernel from the committed ncu capture (% of peak, sustained), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[kernel]["pipes_pct_of_peak"]
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ["index", "clocks.sm", "clocks.max.sm", "clocks_event_reasons.hw_slowdown",
              "clocks_event_reasons.hw_thermal_slowdown", "clocks_event_reasons.sw_thermal_slowdown",
              "clocks_event_reasons.sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = str(gpu_index)
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + ",".join(self.FIELDS),
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == len(self.FIELDS) and parts[0] == self.gpu:
                self.rows.append(parts)

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][2]),
                "reasons": reasons, "samples": len(self.rows)}


def make_host_inputs(n_frames, seed):
    """Synthetic frames on the host: a handful of distinct frames tiled up to the batch (generation is numpy)."""
    from depthinspace_b200 import synth
    base = min(n_frames, 8)
    d = synth.make_frames(base, HW, "default", n_scales=N_SCALES, max_disp=128.0, seed=seed)
    reps = (n_frames + base - 1) // base

    def tile(a):
        return np.ascontiguousarray(np.concatenate([a] * reps, axis=0)[:n_frames])
    return dict(pattern=d["pattern"], im=tile(d["im"]), ambient=tile(d["ambient"]),
                disp=[tile(p) for p in d["disp_pred"]])


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(sample, threads):
    """One pass of the hot path over `sample` frames with the oracle's torch port (the reference's own op
    sequence, model/networks.py + model/ext_functions.py:156-183) on the host cores."""
    from oracle import torch_port
    torch.set_num_threads(threads)
    im = torch.from_numpy(sample["im"])
    amb = torch.from_numpy(sample["ambient"])
    pat_l, _ = torch_port.lcn(torch.from_numpy(sample["pattern"]))
    disps = [torch.from_numpy(p).requires_grad_(True) for p in sample["disp"]]
    t0 = time.perf_counter()
    im_l, im_s = torch_port.lcn(im)
    vals = torch_port.single_frame_loss(disps, im_l, im_s, amb, pat_l, chunk=1)
    total = sum(vals)
    total.backward()
    return time.perf_counter() - t0, float(total.detach())


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 2 if args.steps <= 10 else 1      # bounded sample: keeps K steps within a few minutes of CPU time
    sample = make_host_inputs(n, seed=42)
    for _ in range(args.warmup):
        cpu_reference_step(sample, threads)
    times = [cpu_reference_step(sample, threads)[0] for _ in range(args.steps)]
    t = sum(times)
    value = n * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DIS-SF loss path, {n}-frame sample of the 256-frame batch, {HW[0]}x{HW[1]}, "
                               "default pattern, 4 scales, census_sad 9x9, LCN r5",
                   "note": "reference's CPU implementation = oracle torch port on host cores (pure-Python reference, "
                           "native ext un-vendored)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} frames x {args.steps} steps, all host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch.distributed as dist
    from depthinspace_b200 import _lib, _ops, losses, networks

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    n = args.frames_per_gpu
    host = make_host_inputs(n, seed=42 + rank)
    P = HW[0] * HW[1]
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_im, h_amb = pin(host["im"]), pin(host["ambient"])
    h_disp = [pin(p) for p in host["disp"]]

    lcn = networks.LCN(5, 0.05)
    pat = torch.from_numpy(host["pattern"]).to(dev)
    pat_lcn, _ = lcn(pat)
    loss = losses.SingleFrameLoss(HW[0], HW[1], torch.cat([pat_lcn] * 3, dim=1), process_group=group)

    im, amb = h_im.to(dev), h_amb.to(dev)
    disps = [p.to(dev).requires_grad_(True) for p in h_disp]

    def step(im_d, amb_d, disps_d):
        """LCN + fused loss/gradient kernels: every term of the assembly and d(total)/d(disparity map) for all 4 scales
        (losses.SingleFrameLoss.value_and_grad; the gradients are what autograd.backward(outs, grads) feeds DispNet)."""
        im_l, im_s = lcn(im_d)
        vals, grads = loss.value_and_grad(disps_d, im_l, im_s, amb_d, global_frames=n * world)
        step.grads = grads
        return torch.stack(vals).sum()

    def autograd_step(im_d, amb_d, disps_d):
        """The same through the nn.Module / autograd surface (forward() + backward(), five gradient-scaling passes)."""
        for d in disps_d:
            d.grad = None
        im_l, im_s = lcn(im_d)
        vals = loss(disps_d, im_l, im_s, amb_d)
        total = vals[0]
        for v in vals[1:]:
            total = total + v
        total.backward()
        return total

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # the clock sampler covers warm-up + timed region (the GPU is under the same load throughout), so that
    # even a short timed region gets several 100 ms samples
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            step(im, amb, disps)
        sync_all()

        # ---- device-resident timing (value) ----
        l0 = _lib.LAUNCHES
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            total = step(im, amb, disps)
        ev1.record()
        sync_all()
        launches = _lib.LAUNCHES - l0             # libdis_b200 kernels enqueued inside the timed region
        if ev0.elapsed_time(ev1) < 400.0:      # keep the load on until nvidia-smi has reported at least a few samples
            t_end = time.perf_counter() + 0.5
            while time.perf_counter() < t_end:
                step(im, amb, disps)
            torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    loss_value = float(total.detach())

    # ---- the same work through the nn.Module / autograd surface (forward() + backward()) ----
    for _ in range(2):
        autograd_step(im, amb, disps)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a0.record()
    for _ in range(args.steps):
        a_total = autograd_step(im, amb, disps)
    a1.record()
    torch.cuda.synchronize()
    autograd_ms = a0.elapsed_time(a1) / args.steps
    grad_gap = max(float((g.view_as(d) - d.grad).abs().max() / d.grad.abs().max()) for g, d in zip(step.grads, disps))
    assert abs(float(a_total.detach()) - loss_value) <= 2e-6 * abs(loss_value) and grad_gap < 1e-5, (grad_gap, loss_value)

    # ---- the same step captured once into a CUDA graph and replayed (launch gaps and Python overhead removed) ----
    graph_ms = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            # fresh leaves: autograd binds a leaf's gradient accumulation to the stream of its first use, and the
            # benchmark's own leaves were first used on the legacy default stream, which a capture may not touch
            g_disps = [d.detach().clone().requires_grad_(True) for d in disps]
            for _ in range(2):
                step(im, amb, g_disps)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        for d in g_disps:
            d.grad = None
        with torch.cuda.graph(graph):
            g_total = step(im, amb, g_disps)
        for _ in range(2):
            graph.replay()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(args.steps):
            graph.replay()
        g1.record()
        torch.cuda.synchronize()
        graph_ms = g0.elapsed_time(g1) / args.steps
        assert abs(float(g_total.detach()) - loss_value) <= 1e-6 * abs(loss_value)
        del graph
    except Exception as e:  # graph capture is an optimisation on top of the public API, never a requirement
        import traceback
        traceback.print_exc(file=sys.stderr)
        graph_ms = f"unavailable: {type(e).__name__}: {e}"[:160]

    # ---- dominant kernel alone: 4-scale fused pattern-loss (census_sad 9x9, with gradient stash) ----
    import ctypes
    im_l, im_s = lcn(im)
    gnums = [torch.empty_like(d) for d in disps]
    lib = _lib.load()
    npart = lib.dis_pattern_loss_multi_num_partials(n, HW[0], HW[1])
    partials = torch.empty(2 * N_SCALES * npart, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    PtrArr = ctypes.c_void_p * N_SCALES
    d_arr = PtrArr(*[d.data_ptr() for d in disps])
    g_arr = PtrArr(*[g.data_ptr() for g in gnums])

    def dominant():
        _lib.check(lib.dis_pattern_loss_multi_forward(d_arr, N_SCALES, im_l.data_ptr(), im_s.data_ptr(),
                                                      loss.ph_loss.pattern.data_ptr(), g_arr, partials.data_ptr(),
                                                      n, HW[0], HW[1], 9, 3, 0.5, stream))
    for _ in range(3):
        dominant()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = max(3, min(args.steps, 10))
    k0.record()
    for _ in range(reps):
        dominant()
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / reps

    # ---- end to end: host (pinned) inputs -> public modules -> loss scalar back on the host ----
    # Every step copies ITS OWN inputs (raw IR, ambient, 4 disparity maps = 1.36 GB) from pinned host memory and reads
    # its loss back.  The copies of step i+1 are issued on a side stream while step i computes (double buffering), the
    # way a training loop prefetches; nothing is reused across steps.
    copy_stream = torch.cuda.Stream()

    def upload():
        with torch.cuda.stream(copy_stream):
            bufs = (h_im.to(dev, non_blocking=True), h_amb.to(dev, non_blocking=True),
                    [p.to(dev, non_blocking=True) for p in h_disp])
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs, ev

    def e2e_run(k):
        nxt = upload()
        losses_host = []
        for i in range(k):
            (im_d, amb_d, disps_d), ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            if i + 1 < k:
                nxt = upload()
            for t in (im_d, amb_d, *disps_d):
                t.record_stream(torch.cuda.current_stream())
            total = step(im_d, amb_d, [d.requires_grad_(True) for d in disps_d])
            losses_host.append(float(total.detach()))    # D2H read of the step's loss
        return losses_host

    e2e_run(2)
    sync_all()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    sync_all()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks ----
    times = torch.tensor([ms, e2e_s * 1e3, kernel_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms = times.tolist()
    frames = n * world * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (e2e_ms * 1e-3)
    peak, peak_src = measured_peak()
    achieved = KERNEL_ALGO_BYTES_PER_FRAME * P * n / (kernel_ms * 1e-3) / 1e9
    step_gbs = ALGO_BYTES_PER_FRAME * P * (value / world) / 1e9

    # ---- DIS-MF hot path (BASELINE configs[2]) on this GPU, reported beside the headline (N = 1 only) ----
    mf_line, sfg_line = None, None
    if world == 1 and not args.no_mf:
        try:
            del disps, im, amb
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_mf
            w = bench_mf.build(32, dev)
            for _ in range(3):
                bench_mf.step(w)
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            m0.record()
            for _ in range(args.steps):
                bench_mf.step(w)
            m1.record()
            torch.cuda.synchronize()
            mf_ms = m0.elapsed_time(m1) / args.steps
            mf_line = {"value": 128 / (mf_ms * 1e-3), "unit": UNIT, "ms_per_step": mf_ms, "frames_per_step": 128,
                       "workload": "BASELINE configs[2] on one GPU: DIS-MF hot path, bs 32 x tl 4 (copy_data LCN, 24 xyz/flow + 96 C=32 "
                                   "feature warps fwd/recompute/bwd, 1-scale census_sad + smoothness + 12 flow-consistency terms + L1), "
                                   "see tools/bench_mf.py"}
            del w
            torch.cuda.empty_cache()
            # the DIS-SF assembly INCLUDING its 12 flow-consistency terms (the headline follows SURVEY 8(d): 144P, without)
            w = bench_mf.build_sf(64, dev)
            for _ in range(3):
                bench_mf.step_sf(w)
            torch.cuda.synchronize()
            m0.record()
            for _ in range(args.steps):
                bench_mf.step_sf(w)
            m1.record()
            torch.cuda.synchronize()
            sfg_ms = m0.elapsed_time(m1) / args.steps
            sfg_line = {"value": 256 / (sfg_ms * 1e-3), "unit": UNIT, "ms_per_step": sfg_ms, "frames_per_step": 256,
                        "workload": "DIS-SF loss assembly with its geometric terms (single_frame_worker.py:101-149): copy_data LCN + "
                                    "4 x census_sad 9x9 + smoothness + 6 pairs x 2 directions of the flow-consistency loss, "
                                    "module/autograd path, see tools/bench_mf.py --sf"}
            del w
        except Exception as e:
            if mf_line is None:
                mf_line = f"unavailable: {type(e).__name__}: {e}"[:200]
            else:
                sfg_line = f"unavailable: {type(e).__name__}: {e}"[:200]

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = make_host_inputs(2, seed=42)
            cpu_reference_step(make_host_inputs(1, seed=1), threads)     # warm-up (thread pools, allocator)
            t, _ = cpu_reference_step(sample, threads)
            cpu = {"value": 2 / t, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "2 of the 256 frames, 1 pass, oracle torch port of the reference modules, all host threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: DIS-SF loss path, {n} frames/GPU (bs 64 x tl 4), {HW[0]}x{HW[1]}, "
                                   "default pattern, LCN r5 + 4 x (pattern warp + census_sad 9x9, sigma-weighted) + smoothness, "
                                   "loss terms + gradients w.r.t. all 4 disparity maps (SingleFrameLoss.value_and_grad)",
                       "frames_per_gpu": n, "l2": "inputs (2.7 GB/step) exceed the 126 MB L2; no flush needed",
                       "loss": loss_value},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int((2 + N_SCALES) * 4 * P * n), "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "cuda_graph_ms_per_step": graph_ms, "autograd_modules_ms_per_step": autograd_ms,
            "dis_mf": mf_line, "dis_sf_with_geometric_terms": sfg_line,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic("pattern_multi_kernel<census_sad, R=4, 4 scales, grad>", n),
                         "algorithmic_bytes": KERNEL_ALGO_BYTES_PER_FRAME * P * n,
                         "kernel": "pattern_multi_kernel<census_sad, R=4, 4 scales, grad>",
                         "kernel_ms": kernel_ms, "peak_source": peak_src,
                         "limiter": "XU (rsqrt) + FP32 issue, not HBM: ~100 rsqrt per pixel-scale (see DESIGN.md)",
                         "pipes_pct_of_peak": measured_pipes("pattern_multi_kernel<census_sad, R=4, 4 scales, grad>"),
                         "step_algorithmic_gbs": step_gbs, "step_frac": step_gbs / peak},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=FRAMES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mf", action="store_true", help="skip the DIS-MF side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.warmup < 3:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    run_ours(args)


if __name__ == "__main__":
    main()
