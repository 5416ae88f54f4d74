#!/usr/bin/env python
"""m2d_nvl_allreduce against NCCL on the same data, sizes that exercise every loop tail, plus its time per size:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nvl_check.py
With two ranks the sum is exact in any order, so the comparison is bitwise there; with more ranks the multimem reduction
order inside the switch is not specified and the check is 1e-6 relative."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import dp                                         # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
nvl = dp.NvlAllReduce(dist.group.WORLD, dev)
NMAX = 8 << 20
buf, buf2 = nvl.alloc(NMAX), nvl.alloc(NMAX)
ok = True
g = torch.Generator(device=dev).manual_seed(100 + rank)
for n, off in ((1024, 0), (4 * 513, 4), (4 * (32 * 512 * 4 * world + 3), 8), (4 * (32 * 512 * 7 * world + 511), 0),
               (3_258_752, 16), (7_123_456, 1024)):
    for two in (False, True):
        x = torch.randn(NMAX, device=dev, generator=g)
        y = torch.randn(NMAX, device=dev, generator=g)
        buf.copy_(x)
        buf2.copy_(y)
        ref, ref2 = x.clone(), y.clone()
        dist.all_reduce(ref[off:off + n])
        n2 = (n // 8) * 4
        if two:
            dist.all_reduce(ref2[:n2])
        torch.cuda.synchronize()
        dist.barrier()
        if two:
            nvl.all_reduce_sum2_(buf[off:off + n], buf2[:n2], slot=1)
        else:
            nvl.all_reduce_sum_(buf[off:off + n], slot=0)
        torch.cuda.synchronize()
        nvl.check()
        dist.barrier()
        for got, want, what in ((buf, ref, "segment 0"), (buf2, ref2 if two else y, "segment 1")):
            if world == 2:
                good = torch.equal(got, want)
            else:
                good = bool(((got - want).abs() <= 1e-6 * want.abs().clamp_min(1.0)).all())
            if not good:
                ok = False
                print(f"rank {rank}: MISMATCH n={n} off={off} two={two} {what}: max |d| = {float((got - want).abs().max()):.3e}")
# time per size (all ranks launch back to back; CUDA events on this rank)
for n in (3_258_752, 7_123_456):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        nvl.all_reduce_sum_(buf[:n], slot=0)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(20):
        nvl.all_reduce_sum_(buf[:n], slot=0)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1000
    if rank == 0:
        bus = 2 * (world - 1) / world * n * 4 / (us * 1e-6) / 1e9
        print(f"n = {n} floats ({n * 4 / 1e6:.1f} MB), {nvl.blocks} blocks, multicast {nvl.multicast}: {us:.1f} us, bus bandwidth {bus:.0f} GB/s per GPU")
nvl.check()
print(f"rank {rank}: {'OK' if ok else 'FAILED'}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
