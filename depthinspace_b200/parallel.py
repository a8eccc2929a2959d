"""Data-parallel plumbing for the loss path: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests).

The path shards by sample (a track's frames stay on one rank) and has NO data-path collective.
The only exchanges are
  (1) a handful of fp32 scalars per step: the batch-wide ratios of the reference
      (sum(mask*diff)/sum(mask), model/networks.py:374; means, :431) must be formed from
      all-reduced numerators and denominators, not by averaging per-rank ratios;
  (2) the gradient all-reduce of the (out-of-scope) networks, done on flat buckets.
Because every loss term is normalised by a GLOBAL denominator, per-rank gradients must be SUMMED
(not averaged) across ranks.
"""
import torch
import torch.distributed as dist


def all_reduce_sum_(t, group=None):
    """In-place sum all-reduce of a small tensor; no-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def global_ratio(num_den, group=None):
    """num_den: tensor [..., 2] of per-rank (numerator, denominator) -> global numerator / denominator."""
    nd = num_den.clone()
    all_reduce_sum_(nd, group)
    return nd[..., 0] / nd[..., 1]


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced [begin, end) shard of n_items (samples / tracks) for this rank."""
    base, rem = divmod(n_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_reduce_gradients(params, group=None, bucket_bytes=64 << 20):
    """Sum-all-reduce .grad of params in flat buckets (sized for launch latency, not link count:
    NVSwitch gives every peer full bandwidth)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in grads:
        if bucket and (g.dtype != bucket[0].dtype or size + g.numel() * g.element_size() > bucket_bytes):
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
    flush()
