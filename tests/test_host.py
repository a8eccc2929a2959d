"""Host-side logic on CPU: sharding, the world_size-2 gloo path of the batch-wide ratios and of the flat
gradient all-reduce.  The per-rank partial sums come from the oracle (no GPU here); what is under test is
that all-reducing numerators and denominators reproduces the single-process value of the reference's
ratio (model/networks.py:374), which averaging per-rank ratios would not."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from depthinspace_b200 import parallel, synth


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 129):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synth_is_seeded_and_shaped():
    a = synth.make_frames(2, (32, 40), "kinect", n_scales=3, seed=5)
    b = synth.make_frames(2, (32, 40), "kinect", n_scales=3, seed=5)
    assert a["im"].shape == (2, 1, 32, 40) and a["pattern"].shape == (1, 1, 32, 40) and len(a["disp_pred"]) == 3
    assert all(np.array_equal(a[k], b[k]) for k in ("im", "ambient", "disp_gt", "pattern"))
    assert a["im"].dtype == np.float32 and 0 <= a["im"].min() and a["im"].max() <= 1
    for kind, dens in synth.PATTERN_DENSITY.items():
        assert abs(float(synth.dot_pattern(kind).mean()) - dens) < 0.02


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from oracle import c_oracle
        hw = (24, 32)
        d = synth.make_frames(4, hw, "default", n_scales=1, max_disp=24, seed=9)
        im_l, im_s = c_oracle.lcn_forward(d["im"], 3, 0.05)
        pat_l, _ = c_oracle.lcn_forward(d["pattern"], 3, 0.05)
        b, e = parallel.shard_range(4, rank, ws)
        o = c_oracle.pattern_loss(d["disp_pred"][0][b:e], im_l[b:e], im_s[b:e], pat_l, 5, 3, 0.5, True)
        nd = torch.tensor([o["num"], o["den"]], dtype=torch.float64)
        val = parallel.global_ratio(nd).item()
        # gradient of the GLOBAL ratio w.r.t. local disparities: local d(num)/d(disp) / global den
        local_grad = torch.from_numpy(o["grad_disp"] * o["den"])          # oracle normalised by the local den
        den = nd.clone()
        parallel.all_reduce_sum_(den)
        local_grad = local_grad / den[1]
        # a stand-in "network parameter" whose gradient is the sum of all pixel gradients
        p = torch.nn.Parameter(torch.zeros(3))
        p.grad = torch.stack([local_grad.sum(), local_grad.abs().sum(), torch.tensor(float(rank))]).float()
        parallel.all_reduce_gradients([p])
        q.put((rank, val, p.grad.tolist()))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_reproduces_single_process_ratio():
    from oracle import c_oracle
    hw = (24, 32)
    d = synth.make_frames(4, hw, "default", n_scales=1, max_disp=24, seed=9)
    im_l, im_s = c_oracle.lcn_forward(d["im"], 3, 0.05)
    pat_l, _ = c_oracle.lcn_forward(d["pattern"], 3, 0.05)
    full = c_oracle.pattern_loss(d["disp_pred"][0], im_l, im_s, pat_l, 5, 3, 0.5, True)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, val, grad in res:
        assert abs(val - full["val"]) <= 1e-9 * abs(full["val"])
        assert abs(grad[0] - float(full["grad_disp"].sum())) <= 1e-5 * abs(float(np.abs(full["grad_disp"]).sum()))
        assert abs(grad[1] - float(np.abs(full["grad_disp"]).sum())) <= 1e-5 * float(np.abs(full["grad_disp"]).sum())
        assert grad[2] == 1.0  # 0 + 1: summed, not averaged
    # per-rank ratios averaged would NOT be the reference value (different sigma mass per shard)
    halves = [c_oracle.pattern_loss(d["disp_pred"][0][s], im_l[s], im_s[s], pat_l, 5, 3, 0.5, False)["val"]
              for s in (slice(0, 2), slice(2, 4))]
    assert abs(np.mean(halves) - full["val"]) > 1e-7 * abs(full["val"])
