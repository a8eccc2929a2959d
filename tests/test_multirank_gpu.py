"""Data-parallel semantics on the GPU: a 2-rank run of the loss assembly (every term: photometric x4, smoothness,
geometric pairs, pseudo-GT L1, SGM warm-up) must reproduce the single-process loss and gradients.

reference semantics: the ratios run over the WHOLE batch (model/networks.py:374, :431, :599, :653;
single_frame_worker.py:152-163), so ranks all-reduce numerators and denominators and SUM gradients.

Two ranks are spawned with torch.multiprocessing: NCCL with one GPU per rank when the box has >= 2 GPUs, otherwise both
ranks share cuda:0 and exchange through gloo (NCCL refuses two ranks on one device); the collective code path in
depthinspace_b200 is the same.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HW, TL, BS = (48, 64), 2, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    """Global batch (numpy, seeded): [tl, bs, ...] tensors of the single-frame worker."""
    from depthinspace_b200 import synth
    from oracle import c_oracle
    n = TL * BS
    d = synth.make_frames(n, HW, "default", n_scales=4, max_disp=32, seed=5)
    im_l, im_s = c_oracle.lcn_forward(d["im"], 5, 0.05, "f64")
    pat_l, _ = c_oracle.lcn_forward(d["pattern"], 5, 0.05, "f64")
    rng = np.random.default_rng(11)
    g = synth.make_geometry(n, HW, seed=2)
    flows = {}
    for i in range(TL):
        for j in range(TL):
            if i != j:
                flows[f"flow_{i}{j}"] = synth.make_flows(BS, HW, max_mag=1.5, seed=10 * i + j)[0]
    view = lambda a: a.reshape(TL, BS, *a.shape[1:])
    sgm = (40.0 * rng.random((TL, BS, 1) + HW)).astype(np.float32)        # ~25 % of the pixels above the 30 px threshold
    return dict(
        pattern=np.repeat(pat_l.astype(np.float32), 3, axis=1), K=g["K"],
        im=view(im_l.astype(np.float32)), std=view(im_s.astype(np.float32)), amb=view(d["ambient"]),
        disps=[view((p + 0.2 * rng.standard_normal(p.shape)).astype(np.float32)) for p in d["disp_pred"]],
        pgt=view((d["disp_gt"] + 0.5 * rng.standard_normal(d["disp_gt"].shape)).astype(np.float32)),
        sgm=sgm, noise=[(1.5 * rng.standard_normal(sgm.shape)).astype(np.float32) for _ in range(4)],
        R=view(g["R0"]), t=view(g["t0"]), flows=flows)


def _run(inp, lo, hi, group, device):
    """All terms + gradients of the samples [lo, hi) (batch axis = dim 1), through forward() and value_and_grad()."""
    from depthinspace_b200 import losses
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    sl = lambda a: dev(a[:, lo:hi])
    K = torch.from_numpy(inp["K"].astype(np.float64))
    loss = losses.SingleFrameLoss(HW[0], HW[1], dev(inp["pattern"]), process_group=group, K=K, Ki=torch.linalg.inv(K),
                                  focal_length=float(inp["K"][0, 0]), baseline=0.075)
    outs = [sl(p).requires_grad_(True) for p in inp["disps"]]
    flow_out = {k: dev(v[lo:hi]) for k, v in inp["flows"].items()}
    vals = loss(outs, sl(inp["im"]), sl(inp["std"]), sl(inp["amb"]), pseudo_gt=sl(inp["pgt"]), R=sl(inp["R"]), t=sl(inp["t"]),
                flow_out=flow_out, sgm_disp=sl(inp["sgm"]), sgm_noise=[sl(n) for n in inp["noise"]])
    torch.stack(vals).sum().backward()
    outs2 = [sl(p) for p in inp["disps"]]
    vals2, grads2 = loss.value_and_grad(outs2, sl(inp["im"]), sl(inp["std"]), sl(inp["amb"]), pseudo_gt=sl(inp["pgt"]),
                                        global_frames=TL * BS)
    torch.cuda.synchronize()
    return (torch.stack(vals).detach().cpu().numpy(), [o.grad.cpu().numpy() for o in outs],
            torch.stack(vals2).detach().cpu().numpy(), [g.detach().cpu().numpy() for g in grads2])


def _worker(rank, world, port, backend, out_dir):
    import torch.distributed as dist
    device = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(device)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        inp = _inputs()
        per = BS // world
        res = _run(inp, rank * per, (rank + 1) * per, dist.group.WORLD, device)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), vals=res[0], vals2=res[2],
                 **{f"g{s}": g for s, g in enumerate(res[1])}, **{f"h{s}": g for s, g in enumerate(res[3])})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_reproduce_the_single_process_loss_and_gradients(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    mp.spawn(_worker, args=(world, _free_port(), backend, str(tmp_path)), nprocs=world, join=True)
    inp = _inputs()
    vals, grads, vals2, grads2 = _run(inp, 0, BS, None, torch.device("cuda", 0))
    assert len(vals) == 4 + 1 + 1 + 4 + 4          # photometric, smoothness, 1 frame pair, pseudo-GT x4, SGM x4
    per = BS // world
    worst = {"term": 0.0, "grad": 0.0, "term_vg": 0.0, "grad_vg": 0.0}
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        # every rank holds the GLOBAL value of every term
        worst["term"] = max(worst["term"], float(np.max(np.abs(z["vals"] - vals) / np.maximum(np.abs(vals), 1e-30))))
        worst["term_vg"] = max(worst["term_vg"], float(np.max(np.abs(z["vals2"] - vals2) / np.maximum(np.abs(vals2), 1e-30))))
        for s in range(4):
            ref = grads[s][:, r * per:(r + 1) * per]
            worst["grad"] = max(worst["grad"], float(np.abs(z[f"g{s}"] - ref).max() / np.abs(grads[s]).max()))
            ref2 = grads2[s][:, r * per:(r + 1) * per]
            worst["grad_vg"] = max(worst["grad_vg"], float(np.abs(z[f"h{s}"] - ref2).max() / np.abs(grads2[s]).max()))
    print(f"2-rank ({backend}) vs single process, max relative deviation: {worst}")
    assert worst["term"] <= 1e-6 and worst["term_vg"] <= 1e-6, worst
    assert worst["grad"] <= 1e-6 and worst["grad_vg"] <= 1e-6, worst
