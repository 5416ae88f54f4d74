"""Batch data parallelism for the phase3 train step (the reference has none: SURVEY.md §2.3, §8e).

One process per GPU; every rank runs the full step on its contiguous shard of the global
minibatch.  The critic has no BatchNorm/Dropout (phase3/archis/default.py:249-346) and the
gradient penalty is a mean of per-sample terms (losses.py:56-60), so with equal shards the
global-batch gradient is the mean of the shard gradients: ONE flat all-reduce(sum) per optimizer
step, with the 1/world factor folded into the fused Adam kernel (`gscale`).  Generator BatchNorm
uses per-replica statistics (standard DDP semantics)."""
from __future__ import annotations

import torch


def shard_bounds(global_batch, world, rank):
    """[lo, hi) of rank's contiguous shard; shards must be equal for the mean-of-means identity."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def world_size(group=None):
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size(group)
    return 1


def all_reduce_sum_(flat, group=None):
    """In-place sum of a flat gradient buffer over the ranks (NCCL on GPUs; gloo in the CPU tests).
    Returns the factor the optimizer must apply (1/world)."""
    w = world_size(group)
    if w > 1:
        torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM, group=group)
    return 1.0 / w


def flatten_grads(named_grads, out=None):
    """Concatenate gradients in the given (name, tensor-or-None) order into one flat fp32 buffer
    with 16-byte aligned slots — the layout engine.FlatParams uses for its live parameters."""
    sizes = [(n, 0 if g is None else g.numel()) for n, g in named_grads]
    total = sum((s + 3) // 4 * 4 for _, s in sizes)
    if out is None:
        ref = next(g for _, g in named_grads if g is not None)
        out = torch.zeros(total, dtype=torch.float32, device=ref.device)
    o = 0
    offsets = {}
    for (n, g), (_, s) in zip(named_grads, sizes):
        if g is not None:
            out[o:o + s].copy_(g.reshape(-1))
        offsets[n] = (o, s)
        o += (s + 3) // 4 * 4
    return out, offsets
