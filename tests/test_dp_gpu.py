"""Multi-GPU (needs >= 2 B200s of one node; skipped otherwise): the data-parallel fused train step — one process per
GPU, gradient all-reduce by the NVLink peer-memory kernel (collective="nvl") or by NCCL between per-iteration graphs
(collective="nccl") — against the ORACLE evaluated shard-by-shard with averaged gradients (per-replica
generator BatchNorm statistics, one optimiser step on the mean of the shard gradients), plus the invariant that every
rank ends the step with bit-identical parameters."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
NC, B = 2, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, collective, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {"rank": rank}
    try:
        from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator
        from music2dance_b200.trainer import Phase3Trainer
        from oracle import phase3_oracle as O
        from tests.parity import TOL_NORTH_STAR, scalar_check
        cfg = O.make_cfg(n_critic_steps=NC)
        torch.manual_seed(0)
        gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"],
                                cfg["size"], cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"],
                                cfg["enc_type"], cfg["activ"], dev)
        critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                                       init_ker=cfg["init_kernel"], activ=cfg["activ"], device=dev)
        G0 = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
        D = {k: v.detach().cpu().clone() for k, v in critic.state_dict().items()}
        tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=True, collective=collective)
        res["nvl"] = tr.nvl is not None
        res["multicast"] = bool(tr.nvl.multicast) if tr.nvl is not None else None
        if collective == "nvl":
            assert tr.nvl is not None and not tr.per_iter
        shard = lambda r: [O.synthetic_batch(cfg, B, 6000 + 100 * r + i) for i in range(NC)]
        bs = shard(rank)
        tr.load_batches(*[torch.stack([b[j] for b in bs]) for j in range(4)], bs[-1][4])
        tr.train_step()
        logs = tr.logs()
        if tr.nvl is not None:
            tr.nvl.check()
        # oracle: every shard through the same critic, gradients averaged, ONE Adam step per iteration
        Gs = [{k: v.clone() for k, v in G0.items()} for _ in range(world)]
        ad, ag = O.AdamState(D, cfg["lr_critic"]), O.AdamState(Gs[0], cfg["lr_gen"])
        shards = [shard(r) for r in range(world)]
        for i in range(NC):
            outs = [O.critic_iteration(Gs[r], D, cfg, *shards[r][i][:4], None) for r in range(world)]
            for k in ("loss_critic", "gp", "w_dist"):
                scalar_check(logs["critic"][i][k], outs[rank][k], TOL_NORTH_STAR if i == 0 else 2e-2, f"rank {rank} it{i} {k}")
            avg = {k: (None if outs[0]["grads"][k] is None else sum(o["grads"][k] for o in outs) / world)
                   for k in outs[0]["grads"]}
            with torch.no_grad():
                ad.step(D, avg)
        outs = [O.generator_update(Gs[r], D, cfg, shards[r][-1][0], shards[r][-1][1], shards[r][-1][4], None)
                for r in range(world)]
        for k in ("loss_gen", "l1"):
            scalar_check(logs["gen"][k], outs[rank][k], 2e-2, f"rank {rank} gen {k}")
        avg = {k: (None if outs[0]["grads"][k] is None else sum(o["grads"][k] for o in outs) / world)
               for k in outs[0]["grads"]}
        with torch.no_grad():
            ag.step(Gs[0], avg)
        # parameters after the step: mean deviation from the oracle in units of lr (tests/parity.py TOL_DRIFT_LR)
        worst = 0.0
        for mod, P, lr, steps in ((critic, D, cfg["lr_critic"], NC), (gen, Gs[0], cfg["lr_gen"], 1)):
            skip = set(O.pre_bn_bias_names(P))
            for k, v in mod.state_dict().items():
                if k in skip or not v.is_floating_point() or "running_" in k:
                    continue
                d = float((v.cpu() - P[k]).abs().mean()) / (lr * steps)
                worst = max(worst, d)
                assert d < 0.25 + 1e-6 * float(P[k].abs().max()) / (lr * steps), (k, d)
        res["worst_drift_lr"] = worst
        # every replica holds the same parameters (the all-reduce gave every rank the same sums)
        for name, flat in (("gen", tr.ge.fp.flat), ("critic", tr.de.fp.flat)):
            ref = flat.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref, flat), f"rank {rank}: {name} parameters differ from rank 0 after the step"
        res["ok"] = True
    except Exception as e:          # noqa: BLE001
        import traceback
        res["ok"], res["err"] = False, traceback.format_exc()[-1500:]
    q.put(res)
    try:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:               # noqa: BLE001
        pass
    os._exit(0)


@pytest.mark.parametrize("collective", ["nvl", "nccl"])
def test_two_gpu_step_vs_sharded_oracle(collective):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, collective, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for r in res:
        assert r["ok"], f"rank {r['rank']} ({collective}): {r.get('err')}"
    print(collective, res)
