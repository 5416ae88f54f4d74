"""phase2 unconditional sequence WGAN-LP (BASELINE.json configs[1]): the CPU oracle against values
produced by the reference's own phase2/archis/default.py + losses (tests/golden/phase2.npz), and (GPU) the CUDA
path (music2dance_b200/phase2.py) against the same fixture and the oracle.
Tolerances as in tests/parity.py: scalars 2e-4 (north star 1e-3), gradient digests statistical (ReLU kinks)."""
import os

import numpy as np
import pytest
import torch

from oracle import phase2_oracle as P2
from oracle import phase3_oracle as O
from tests.parity import TOL_FP32, TOL_GEN_GRAD_E2E, TOL_GRAD, TOL_GRAD_BIAS, digest_check, scalar_check

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phase2.npz")
SEED_STEP, SEED_DATA = 31, 78


def oracle_state(cfg, state):
    torch.manual_seed(0)
    G, D = P2.init_generator_params(cfg), P2.init_critic_params(cfg)
    if state == "perturbed":
        O.perturb_params(G)
        O.perturb_params(D)
    return G, D


def pre_bn_biases(cfg):
    return {"decoder.fc1.bias"} | {f"decoder.blocks.{b}.fc2.bias" for b in range(cfg["nblocks_gen"])}


@pytest.mark.parametrize("state", ["init", "perturbed"])
def test_oracle_matches_reference_phase2(state):
    torch.set_num_threads(8)
    gold = np.load(GOLD)
    cfg = P2.make_cfg()
    B, T = int(gold["B"]), cfg["stick_length"]
    G, D = oracle_state(cfg, state)
    for k, v in list(G.items()) + [("D." + k, v) for k, v in D.items()]:
        digest_check(v, gold, f"{state}/init/{k}", 1e-7, f"init {k}")
    real_bt = P2.synthetic_poses(B, T, SEED_DATA)
    torch.manual_seed(SEED_STEP)
    noise = torch.randn(B, T, cfg["input_vector_size"])
    alpha = torch.rand(B, 1)
    o = P2.critic_iteration(G, D, cfg, real_bt, noise, alpha)
    for k in ("loss_critic", "gp", "w_dist"):
        scalar_check(o[k], gold[f"{state}/critic/{k}"], TOL_FP32, k)
    ref_fake = torch.from_numpy(gold[f"{state}/critic/fake"])
    assert float((o["fake"] - ref_fake).abs().max()) < TOL_FP32 * float(ref_fake.abs().max())
    for k, g in o["grads"].items():
        digest_check(g, gold, f"{state}/critic/grad/{k}", TOL_GRAD, f"critic grad {k}", abs_floor=1e-4, kinks=True)
    noise = torch.randn(B, T, cfg["input_vector_size"])
    o = P2.generator_update(G, D, cfg, real_bt, noise)
    scalar_check(o["loss_gen"], gold[f"{state}/gen/loss_gen"], TOL_FP32, "loss_gen")
    scalar_check(o["tv"], gold[f"{state}/gen/tv"], TOL_FP32, "tv")
    skip = pre_bn_biases(cfg)
    for k, g in o["grads"].items():
        if g is None:
            assert f"{state}/gen/nograd/{k}" in gold.files, k
        elif k not in skip:
            digest_check(g, gold, f"{state}/gen/grad/{k}", TOL_GEN_GRAD_E2E, f"gen grad {k}", abs_floor=1e-4, kinks=True)
    for k, v in G.items():
        if "running" in k or "num_batches" in k:
            digest_check(v, gold, f"{state}/gen/buf/{k}", TOL_FP32, f"bn buffer {k}")


def _build(cfg, state, dev):
    from music2dance_b200.phase2 import SequenceDiscriminator, SequenceGenerator
    torch.manual_seed(0)
    gen = SequenceGenerator(cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"], cfg["output_size"],
                            cfg["nblocks_gen"], cfg["n_cells"], dev)
    critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["stick_length"],
                                   init_ker=cfg["init_kernel"], n_blocks=cfg["nblocks_critic"], device=dev)
    if state == "perturbed":
        for m in (gen, critic):
            sd = {k: v.cpu() for k, v in m.state_dict().items()}
            O.perturb_params(sd)
            m.load_state_dict(sd, strict=True)
    return gen, critic


@pytest.mark.gpu
@pytest.mark.parametrize("state", ["init", "perturbed"])
def test_cuda_phase2_step_vs_reference_fixture(state):
    from music2dance_b200.phase2 import Phase2Trainer
    dev = "cuda:0"
    gold = np.load(GOLD)
    cfg = P2.make_cfg()
    B, T = int(gold["B"]), cfg["stick_length"]
    gen, critic = _build(cfg, state, dev)
    for k, v in list(gen.state_dict().items()) + [("D." + k, v) for k, v in critic.state_dict().items()]:
        digest_check(v, gold, f"{state}/init/{k}", 1e-7, f"init {k}")          # same keys, same initial weights
    tr = Phase2Trainer(gen, critic, cfg, B)
    real_bt = P2.synthetic_poses(B, T, SEED_DATA)
    torch.manual_seed(SEED_STEP)
    noise = torch.randn(B, T, cfg["input_vector_size"])
    alpha = torch.rand(B, 1)
    logs = tr.critic_iteration(real_bt, noise, alpha, update=False)
    for k in ("loss_critic", "gp", "w_dist"):
        scalar_check(logs[k], gold[f"{state}/critic/{k}"], TOL_FP32, k)
    ref_fake = torch.from_numpy(gold[f"{state}/critic/fake"])
    fake = tr.fake.view(B, T, cfg["output_size"]).permute(0, 2, 1).cpu()
    assert float((fake - ref_fake).abs().max()) < TOL_FP32 * float(ref_fake.abs().max())
    for k, g in tr.critic_grads().items():
        digest_check(g, gold, f"{state}/critic/grad/{k}", TOL_GRAD_BIAS if k.endswith(".bias") else TOL_GRAD,
                     f"critic grad {k}", abs_floor=1e-4, kinks=True)
    noise = torch.randn(B, T, cfg["input_vector_size"])
    logs = tr.generator_update(real_bt, noise, update=False)
    scalar_check(logs["loss_gen"], gold[f"{state}/gen/loss_gen"], TOL_FP32, "loss_gen")
    scalar_check(logs["tv"], gold[f"{state}/gen/tv"], TOL_FP32, "tv")
    skip = pre_bn_biases(cfg)
    for k, g in tr.generator_grads().items():
        if f"{state}/gen/nograd/{k}" in gold.files:
            assert float(g.abs().max()) == 0.0, k
        elif k not in skip:
            digest_check(g, gold, f"{state}/gen/grad/{k}", TOL_GEN_GRAD_E2E, f"gen grad {k}", abs_floor=1e-4, kinks=True)
    for k, v in gen.state_dict().items():
        if "running" in k or "num_batches" in k:
            digest_check(v, gold, f"{state}/gen/buf/{k}", TOL_FP32, f"bn buffer {k}")


@pytest.mark.gpu
def test_cuda_phase2_batch24_vs_oracle():
    """default.yaml batch (24): one critic iteration + generator update with Adam against the oracle."""
    from music2dance_b200.phase2 import Phase2Trainer
    dev = "cuda:0"
    cfg = P2.make_cfg()
    B, T = cfg["batch_size"], cfg["stick_length"]
    gen, critic = _build(cfg, "perturbed", dev)
    G = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    D = {k: v.detach().cpu().clone() for k, v in critic.state_dict().items()}
    tr = Phase2Trainer(gen, critic, cfg, B)
    real_bt = P2.synthetic_poses(B, T, 5)
    torch.manual_seed(9)
    noise, alpha = torch.randn(B, T, cfg["input_vector_size"]), torch.rand(B, 1)
    o = P2.critic_iteration(G, D, cfg, real_bt, noise, alpha)
    logs = tr.critic_iteration(real_bt, noise, alpha, update=True)
    for k in ("loss_critic", "gp", "w_dist"):
        scalar_check(logs[k], o[k], 1e-3, k)
    noise = torch.randn(B, T, cfg["input_vector_size"])
    # oracle critic after one Adam step (lr 5e-4, first step = -lr * sign-like update)
    st = O.AdamState.__new__(O.AdamState)
    st.lr, st.b1, st.b2, st.eps, st.t = cfg["lr_critic"], 0.9, 0.999, 1e-8, {}
    st.m = {k: torch.zeros_like(D[k]) for k in P2.trainable(D)}
    st.v = {k: torch.zeros_like(D[k]) for k in P2.trainable(D)}
    with torch.no_grad():
        st.step(D, o["grads"])
    o = P2.generator_update(G, D, cfg, real_bt, noise)
    logs = tr.generator_update(real_bt, noise, update=True)
    scalar_check(logs["loss_gen"], o["loss_gen"], 2e-2, "loss_gen after one critic Adam step")
    scalar_check(logs["tv"], o["tv"], 1e-3, "tv")


def test_lr_schedule_matches_multisteplr():
    """phase2/train.py:88-89,179-180: MultiStepLR(milestones, gamma=0.8) stepped once per generator update.  Host-only:
    the trainer's factor against torch's scheduler (milestones scaled down so that the loop is short)."""
    from music2dance_b200.phase2 import Phase2Trainer

    class T(Phase2Trainer):
        LR_MILESTONES = (10, 35, 50)

        def __init__(self):              # no device needed for the schedule
            self.sched_steps = 0

    t = T()
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=2e-4)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[10, 35, 50], gamma=0.8)
    for _ in range(60):
        assert abs(2e-4 * t.lr_factor() - opt.param_groups[0]["lr"]) < 1e-12, t.sched_steps
        opt.step()
        sch.step()
        t.sched_steps += 1
    assert Phase2Trainer.LR_MILESTONES == (10000, 35000, 50000) and Phase2Trainer.LR_GAMMA == 0.8
