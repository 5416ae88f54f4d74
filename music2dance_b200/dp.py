"""Batch data parallelism for the phase3 train step (the reference has none: SURVEY.md §2.3, §8e).

One process per GPU; every rank runs the full step on its contiguous shard of the global
minibatch.  The critic has no BatchNorm/Dropout (phase3/archis/default.py:249-346) and the
gradient penalty is a mean of per-sample terms (losses.py:56-60), so with equal shards the
global-batch gradient is the mean of the shard gradients: ONE flat all-reduce(sum) per optimizer
step, with the 1/world factor folded into the fused Adam kernel (`gscale`).  Generator BatchNorm
uses per-replica statistics (standard DDP semantics)."""
from __future__ import annotations

import os

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


class NvlAllReduce:
    """Gradient all-reduce over NVLink / NVSwitch peer memory: hand-written kernel `m2d_nvl_allreduce` (two-shot,
    NVLink-SHARP multimem.ld_reduce / multimem.st when the buffers have a multicast mapping, peer loads / stores
    otherwise) on buffers allocated from torch's symmetric-memory allocator.  It is an ordinary kernel launch on the
    caller's stream, so it is captured into the train step's CUDA graph like everything else — no NCCL call, no
    graph split at the collective.  One instance per process; `alloc` must be called in the same order on every rank."""

    def __init__(self, group=None, device=None, blocks=32, pad_bytes=16384):
        import ctypes as C
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self._C, self._symm = C, symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.blocks = int(os.environ.get("M2D_NVL_BLOCKS", blocks))
        if symm.get_signal_pad_size() < pad_bytes:
            symm.set_signal_pad_size(pad_bytes)
        self.pad_slots = symm.get_signal_pad_size() // 4
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.regs = []                     # (tensor, handle, peer pointer array, pad pointer array, multicast pointer)
        self.multicast = None

    def alloc(self, n):
        """Zeroed symmetric float buffer of n floats, mapped on every rank (collective call)."""
        C = self._C
        t = self._symm.empty(n, dtype=torch.float32, device=self.device)
        hdl = self._symm.rendezvous(t, self.group)
        t.zero_()
        ptrs = (C.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        pads = (C.c_void_p * self.world)(*[int(p) for p in hdl.signal_pad_ptrs])
        mc = int(hdl.multicast_ptr) if getattr(hdl, "has_multicast_support", False) and hdl.multicast_ptr else 0
        assert int(ptrs[self.rank]) == t.data_ptr(), "symmetric-memory handle does not map the local tensor at offset 0"
        self.regs.append((t, hdl, ptrs, pads, mc))
        self.multicast = bool(mc) if self.multicast is None else (self.multicast and bool(mc))
        return t

    def _reg_of(self, t):
        p = t.data_ptr()
        for r in self.regs:
            base = r[0].data_ptr()
            if base <= p < base + 4 * r[0].numel():
                return r, (p - base) // 4
        raise ValueError("tensor is not a view of a buffer from NvlAllReduce.alloc")

    def all_reduce_sum_(self, t, slot=0, blocks=None):
        """In-place sum over the ranks of the flat float view `t` (numel and offset multiples of 4).  `slot`: distinct
        per call site that may be in flight concurrently with another one on the same buffer (flag-array region)."""
        from . import ops
        (buf, _, ptrs, pads, mc), off = self._reg_of(t)
        nb = blocks or self.blocks
        slot0 = slot * self.blocks * self.world
        assert slot0 + nb * self.world <= self.pad_slots, "signal pad too small for this many concurrent call sites"
        n = t.numel()
        assert n % 4 == 0 and off % 4 == 0 and t.is_contiguous()
        ops.nvl_allreduce(ptrs, mc, pads, self.rank, self.world, off, n, nb, slot0, self.status)

    def all_reduce_sum2_(self, t, t2, slot=0, blocks=None):
        """Both flat views in ONE launch (one pair of handshakes); the flag array of `t`'s buffer is used."""
        from . import ops
        (_, _, ptrs, pads, mc), off = self._reg_of(t)
        (_, _, ptrs2, _, mc2), off2 = self._reg_of(t2)
        nb = blocks or self.blocks
        slot0 = slot * self.blocks * self.world
        assert slot0 + nb * self.world <= self.pad_slots
        for x, o in ((t, off), (t2, off2)):
            assert x.numel() % 4 == 0 and o % 4 == 0 and x.is_contiguous()
        ops.nvl_allreduce2(ptrs, mc, off, t.numel(), ptrs2, mc2, off2, t2.numel(), pads, self.rank, self.world, nb,
                           slot0, self.status)

    def check(self):
        """Raise if any launch so far timed out waiting for a peer (synchronises)."""
        if int(self.status.item()) != 0:
            raise RuntimeError("m2d_nvl_allreduce: a peer did not reach the collective within 2 s")
