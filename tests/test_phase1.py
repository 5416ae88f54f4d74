"""phase1 stick-figure WGAN-GP (BASELINE.json configs[0]): the CPU oracle against values produced
by the reference's own phase1/archis/residual.py + losses.gradient_penalty (tests/golden/phase1.npz), and (GPU)
the CUDA path (music2dance_b200/phase1.py) against the oracle and the same fixture.
Tolerances: scalars 2e-4 relative (north star 1e-3); gradient digests as in tests/parity.py (ReLU kinks)."""
import os

import numpy as np
import pytest
import torch

from oracle import phase1_oracle as P1
from oracle import phase3_oracle as O
from tests.parity import TOL_FP32, TOL_GRAD, TOL_GRAD_BIAS, digest_check, scalar_check

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phase1.npz")
CONFIGS = {"b2l50s256": {}, "b1l10s32": dict(nblocks_gen=1, nblocks_critic=1, latent_vector_size=10, size=32)}
SEED_STEP, SEED_DATA = 2024, 77


def oracle_state(cfg):
    torch.manual_seed(0)
    G, D = P1.init_generator_params(cfg), P1.init_critic_params(cfg)
    O.perturb_params(G)
    O.perturb_params(D)
    return G, D


@pytest.mark.parametrize("name", list(CONFIGS))
def test_oracle_matches_reference_phase1(name):
    gold = np.load(GOLD)
    cfg = P1.make_cfg(**CONFIGS[name])
    B = cfg["batch_size"]
    G, D = oracle_state(cfg)
    for k, v in list(G.items()) + [("D." + k, v) for k, v in D.items()]:
        digest_check(v, gold, f"{name}/init/{k}", 1e-7, f"init {k}")
    real = P1.synthetic_poses(B, SEED_DATA)
    torch.manual_seed(SEED_STEP)
    noise, mask_g, alpha, masks_d = P1.draw_critic_randoms(cfg, B)
    o = P1.critic_iteration(G, D, cfg, real, noise, mask_g, alpha, masks_d)
    for k in ("loss_critic", "gp", "w_dist"):
        scalar_check(o[k], gold[f"{name}/critic/{k}"], 1e-5, k)
    ref_fake = torch.from_numpy(gold[f"{name}/critic/fake"])
    assert float((o["fake"] - ref_fake).abs().max()) < 1e-5 * float(ref_fake.abs().max())
    for k, g in o["grads"].items():
        if g is None:
            assert f"{name}/critic/nograd/{k}" in gold.files, k
        else:
            digest_check(g, gold, f"{name}/critic/grad/{k}", 1e-4, f"critic grad {k}", abs_floor=1e-5)
    noise, mask_g, masks_d = P1.draw_gen_randoms(cfg, B)
    o = P1.generator_update(G, D, cfg, real, noise, mask_g, masks_d)
    scalar_check(o["loss_gen"], gold[f"{name}/gen/loss_gen"], 1e-5, "loss_gen")
    # biases in front of a train-mode BatchNorm have an exactly-zero true gradient: what any implementation
    # returns there is summation noise (1e-9), different on every host -> not compared
    skip = {"fc1.bias"} | {f"blocks.{i}.fc2.bias" for i in range(cfg["nblocks_gen"])}
    for k, g in o["grads"].items():
        if g is None:
            assert f"{name}/gen/nograd/{k}" in gold.files, k
        elif k not in skip:
            digest_check(g, gold, f"{name}/gen/grad/{k}", 1e-4, f"gen grad {k}", abs_floor=1e-5)
    for k, v in G.items():
        if "running" in k or "num_batches" in k:
            digest_check(v, gold, f"{name}/gen/buf/{k}", 1e-5, f"bn buffer {k}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CONFIGS))
def test_cuda_phase1_step_vs_reference_and_oracle(name):
    from music2dance_b200.phase1 import Discriminator, Generator, Phase1Trainer
    dev = "cuda:0"
    gold = np.load(GOLD)
    cfg = P1.make_cfg(**CONFIGS[name])
    B = cfg["batch_size"]
    torch.manual_seed(0)
    gen = Generator(cfg["latent_vector_size"], cfg["size"], cfg["output_size"], cfg["nblocks_gen"]).to(dev)
    critic = Discriminator(cfg["output_size"], cfg["size"], cfg["nblocks_critic"]).to(dev)
    for m in (gen, critic):
        sd = {k: v.cpu() for k, v in m.state_dict().items()}
        O.perturb_params(sd)
        m.load_state_dict(sd, strict=True)
    for k, v in list(gen.state_dict().items()) + [("D." + k, v) for k, v in critic.state_dict().items()]:
        digest_check(v, gold, f"{name}/init/{k}", 1e-7, f"init {k}")          # same keys, same default init
    tr = Phase1Trainer(gen, critic, cfg, B)
    real = P1.synthetic_poses(B, SEED_DATA)
    torch.manual_seed(SEED_STEP)
    noise, mask_g, alpha, masks_d = P1.draw_critic_randoms(cfg, B)
    logs = tr.critic_iteration(real, noise, mask_g, alpha, masks_d, update=False)
    for k in ("loss_critic", "gp", "w_dist"):
        scalar_check(logs[k], gold[f"{name}/critic/{k}"], TOL_FP32, k)
    ref_fake = torch.from_numpy(gold[f"{name}/critic/fake"])
    assert float((tr.fake.cpu() - ref_fake).abs().max()) < TOL_FP32 * float(ref_fake.abs().max())
    for k, g in tr.critic_grads().items():
        if f"{name}/critic/nograd/{k}" in gold.files:
            assert float(g.abs().max()) == 0.0, k                              # dead fc1 branch (Q1): never touched
        else:
            digest_check(g, gold, f"{name}/critic/grad/{k}", TOL_GRAD_BIAS if k.endswith(".bias") else TOL_GRAD,
                         f"critic grad {k}", abs_floor=1e-4, kinks=True)
    noise, mask_g, masks_d = P1.draw_gen_randoms(cfg, B)
    logs = tr.generator_update(real, noise, mask_g, masks_d, update=False)
    scalar_check(logs["loss_gen"], gold[f"{name}/gen/loss_gen"], TOL_FP32, "loss_gen")
    skip = {"fc1.bias"} | {f"blocks.{i}.fc2.bias" for i in range(cfg["nblocks_gen"])}   # pre-BN biases: exact zero gradient
    for k, g in tr.generator_grads().items():
        if f"{name}/gen/nograd/{k}" in gold.files:
            assert float(g.abs().max()) == 0.0, k
        elif k not in skip:
            digest_check(g, gold, f"{name}/gen/grad/{k}", TOL_GRAD, f"gen grad {k}", abs_floor=1e-4, kinks=True)
    for k, v in gen.state_dict().items():
        if "running" in k or "num_batches" in k:
            digest_check(v, gold, f"{name}/gen/buf/{k}", TOL_FP32, f"bn buffer {k}")


@pytest.mark.gpu
def test_cuda_phase1_chained_steps_vs_oracle():
    """Two full train steps (n_critic critic iterations + generator update, Adam included) against the oracle."""
    from music2dance_b200.phase1 import Discriminator, Generator, Phase1Trainer
    dev = "cuda:0"
    cfg = P1.make_cfg(n_critic_steps=2)
    B = cfg["batch_size"]
    torch.manual_seed(0)
    gen = Generator(cfg["latent_vector_size"], cfg["size"], cfg["output_size"], cfg["nblocks_gen"]).to(dev)
    critic = Discriminator(cfg["output_size"], cfg["size"], cfg["nblocks_critic"]).to(dev)
    G = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    D = {k: v.detach().cpu().clone() for k, v in critic.state_dict().items()}
    tr = Phase1Trainer(gen, critic, cfg, B)
    ad, ag = O.AdamState.__new__(O.AdamState), O.AdamState.__new__(O.AdamState)
    for st, P, lr in ((ad, D, cfg["lr_critic"]), (ag, G, cfg["lr_gen"])):
        st.lr, st.b1, st.b2, st.eps, st.t = lr, 0.9, 0.999, 1e-8, {}
        st.m = {k: torch.zeros_like(P[k]) for k in P1.trainable(P)}
        st.v = {k: torch.zeros_like(P[k]) for k in P1.trainable(P)}
    torch.manual_seed(5)
    it = 0
    for step in range(2):
        for i in range(cfg["n_critic_steps"]):
            real = P1.synthetic_poses(B, 900 + it)
            it += 1
            r = P1.draw_critic_randoms(cfg, B)
            o = P1.critic_iteration(G, D, cfg, real, *r)
            with torch.no_grad():
                ad.step(D, o["grads"])
            logs = tr.critic_iteration(real, *r, update=True)
            for k in ("loss_critic", "gp", "w_dist"):
                scalar_check(logs[k], o[k], 1e-3 if it == 1 else 2e-2, f"step{step} it{i} {k}")
        r = P1.draw_gen_randoms(cfg, B)
        o = P1.generator_update(G, D, cfg, real, *r)
        with torch.no_grad():
            ag.step(G, o["grads"])
        logs = tr.generator_update(real, *r, update=True)
        scalar_check(logs["loss_gen"], o["loss_gen"], 2e-2, f"step{step} loss_gen")
