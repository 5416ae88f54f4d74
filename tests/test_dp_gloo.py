"""CPU, world_size 2, gloo: the data-parallel host logic (music2dance_b200/dp.py) — shard bounds and
the flat gradient all-reduce — checked with oracle gradients: all-reduce(sum)/world of the shard
gradients equals the global-batch critic gradient (no BatchNorm in the critic; GP is a per-sample mean)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from music2dance_b200 import dp
        from oracle import phase3_oracle as O
        cfg = O.make_cfg(ablated=True)
        torch.manual_seed(0)
        D = O.init_critic_params(cfg)
        Bg = 4
        real, audio, noise, alpha, _ = O.synthetic_batch(cfg, Bg, 31)
        g = torch.Generator().manual_seed(5)
        fake = torch.rand(Bg, cfg["output_size"], cfg["stick_length"], generator=g)      # same "generated" poses everywhere

        def critic_grads(lo, hi):
            Dl = O._leaf({k: v.clone() for k, v in D.items()})
            r = real[lo:hi].reshape(hi - lo, cfg["stick_length"], cfg["output_size"]).permute(0, 2, 1).contiguous()
            gp, _, _ = O.gradient_penalty(Dl, cfg, r, fake[lo:hi], None, alpha[lo:hi])
            err = O.critic_forward(Dl, cfg, fake[lo:hi]).mean() - O.critic_forward(Dl, cfg, r).mean() + cfg["gamma"] * gp
            names = O.trainable_names(D)
            gl = torch.autograd.grad(err, [Dl[k] for k in names], allow_unused=True)
            return list(zip(names, gl))

        lo, hi = dp.shard_bounds(Bg, world, rank)
        assert (lo, hi) == (rank * 2, rank * 2 + 2)
        flat, offs = dp.flatten_grads(critic_grads(lo, hi))
        scale = dp.all_reduce_sum_(flat)
        assert scale == 1.0 / world and dp.world_size() == world
        flat *= scale
        full, offs2 = dp.flatten_grads(critic_grads(0, Bg))
        assert offs == offs2
        err = float((flat - full).abs().max() / full.abs().max())
        # every rank holds the same reduced buffer
        chk = flat.clone()
        dist.broadcast(chk, 0)
        same = bool(torch.equal(chk, flat))
        with pytest.raises(ValueError):
            dp.shard_bounds(5, world, rank)
        q.put((rank, err, same))
    finally:
        dist.destroy_process_group()


def test_dp_allreduce_equals_global_batch_gradient():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same in res:
        assert same, f"rank {rank}: reduced buffers differ between ranks"
        # fp32 summation-order noise only; ReLU kinks can move isolated entries (tests/parity.py)
        assert err < 1e-3, f"rank {rank}: shard-mean gradient deviates from the global-batch gradient by {err:.3e}"
