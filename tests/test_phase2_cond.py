"""phase2 conditional (dance-type label) sequence WGAN (BASELINE.json configs[2]): the CPU oracle
against values produced by the reference's own phase2/archis/conditional.py + losses (tests/golden/phase2_cond.npz),
and (GPU) the CUDA path (music2dance_b200/phase2_cond.py) against the same fixture and the oracle.
Tolerances as in tests/parity.py: scalars 2e-4 (north star 1e-3), gradient digests statistical (ReLU kinks)."""
import os

import numpy as np
import pytest
import torch

from oracle import phase2_cond_oracle as PC
from oracle import phase3_oracle as O
from tests.parity import TOL_FP32, TOL_GEN_GRAD_E2E, TOL_GRAD, TOL_GRAD_BIAS, digest_check, scalar_check

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phase2_cond.npz")
SEED_STEP, SEED_DATA, SEED_LABEL = 47, 91, 5


def oracle_state(cfg, state):
    torch.manual_seed(0)
    G, D = PC.init_generator_params(cfg), PC.init_critic_params(cfg)
    if state == "perturbed":
        O.perturb_params(G)
        O.perturb_params(D)
    return G, D


def pre_bn_biases(cfg):
    return {"decoder.fc1.bias"} | {f"decoder.blocks.{b}.fc2.bias" for b in range(cfg["nblocks_gen"])}


@pytest.mark.parametrize("state", ["init", "perturbed"])
def test_oracle_matches_reference_phase2_cond(state):
    torch.set_num_threads(8)
    gold = np.load(GOLD)
    cfg = PC.make_cfg()
    B, T = int(gold["B"]), cfg["stick_length"]
    G, D = oracle_state(cfg, state)
    for k, v in list(G.items()) + [("D." + k, v) for k, v in D.items()]:
        digest_check(v, gold, f"{state}/init/{k}", 1e-7, f"init {k}")
    real_bt = PC.synthetic_poses(B, T, SEED_DATA)
    labels = PC.synthetic_labels(B, SEED_LABEL)
    assert np.array_equal(labels.numpy(), gold["labels"])
    torch.manual_seed(SEED_STEP)
    o = PC.critic_iteration(G, D, cfg, real_bt, labels, *PC.draw_critic_inputs(cfg, B))
    for k in ("loss_critic", "gp", "w_dist"):
        scalar_check(o[k], gold[f"{state}/critic/{k}"], TOL_FP32, k)
    ref_fake = torch.from_numpy(gold[f"{state}/critic/fake"])
    assert float((o["fake"] - ref_fake).abs().max()) < TOL_FP32 * float(ref_fake.abs().max())
    for k, g in o["grads"].items():
        digest_check(g, gold, f"{state}/critic/grad/{k}", TOL_GRAD, f"critic grad {k}", abs_floor=1e-4, kinks=True)
    o = PC.generator_update(G, D, cfg, real_bt, labels, *PC.draw_generator_inputs(cfg, B))
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
    # eval mode: running statistics, no dropout
    real = real_bt.reshape(B, T, cfg["output_size"]).permute(0, 2, 1).contiguous()
    noise = torch.randn(B, T, cfg["input_vector_size"], generator=torch.Generator().manual_seed(3))
    fake = PC.generator_forward(G, cfg, noise, labels, None, train=False)
    ref = torch.from_numpy(gold[f"{state}/eval/fake"])
    assert float((fake - ref).abs().max()) < TOL_FP32 * float(ref.abs().max())
    score = PC.critic_forward(D, cfg, real, labels, None)
    ref = torch.from_numpy(gold[f"{state}/eval/score"])
    assert float((score - ref).abs().max()) < TOL_FP32 * max(1.0, float(ref.abs().max()))


def _build(cfg, state, dev):
    from music2dance_b200.phase2_cond import SequenceDiscriminator, SequenceGenerator
    torch.manual_seed(0)
    gen = SequenceGenerator(cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"], cfg["output_size"],
                            cfg["nblocks_gen"], cfg["n_cells"]).to(dev)
    critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["stick_length"],
                                   init_ker=cfg["init_kernel"], n_blocks=cfg["nblocks_critic"]).to(dev)
    if state == "perturbed":
        for m in (gen, critic):
            sd = {k: v.cpu() for k, v in m.state_dict().items()}
            O.perturb_params(sd)
            m.load_state_dict(sd, strict=True)
    return gen, critic


def test_dropin_constructors_match_reference_fixture_on_cpu():
    """Same state_dict keys and the same initial weights under the same seed (constructor RNG order incl. the
    Embedding draws) — checked without a GPU; forward on CPU tensors must refuse (no fallback)."""
    gold = np.load(GOLD)
    cfg = PC.make_cfg()
    gen, critic = _build(cfg, "init", "cpu")
    sds = list(gen.state_dict().items()) + [("D." + k, v) for k, v in critic.state_dict().items()]
    ref_keys = [k[len("init/init/"):-len("/sum")] for k in gold.files if k.startswith("init/init/") and k.endswith("/sum")]
    assert [k for k, _ in sds] == ref_keys
    for k, v in sds:
        digest_check(v, gold, f"init/init/{k}", 1e-7, f"init {k}")
    with pytest.raises(RuntimeError):
        gen(torch.zeros(2, 120, 50), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        critic(torch.zeros(2, 69, 120), torch.zeros(2, dtype=torch.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("state", ["init", "perturbed"])
def test_cuda_phase2_cond_step_vs_reference_fixture(state):
    from music2dance_b200.phase2_cond import Phase2CondTrainer
    dev = "cuda:0"
    gold = np.load(GOLD)
    cfg = PC.make_cfg()
    B, T = int(gold["B"]), cfg["stick_length"]
    gen, critic = _build(cfg, state, dev)
    for k, v in list(gen.state_dict().items()) + [("D." + k, v) for k, v in critic.state_dict().items()]:
        digest_check(v, gold, f"{state}/init/{k}", 1e-7, f"init {k}")
    tr = Phase2CondTrainer(gen, critic, cfg, B)
    real_bt = PC.synthetic_poses(B, T, SEED_DATA)
    labels = PC.synthetic_labels(B, SEED_LABEL)
    torch.manual_seed(SEED_STEP)
    noise, mask_g, alpha, masks_d = PC.draw_critic_inputs(cfg, B)
    logs = tr.critic_iteration(real_bt, labels, noise, mask_g, alpha, masks_d, update=False)
    for k in ("loss_critic", "gp", "w_dist"):
        scalar_check(logs[k], gold[f"{state}/critic/{k}"], TOL_FP32, k)
    ref_fake = torch.from_numpy(gold[f"{state}/critic/fake"])
    fake = tr.fake.view(B, T, cfg["output_size"]).permute(0, 2, 1).cpu()
    assert float((fake - ref_fake).abs().max()) < TOL_FP32 * float(ref_fake.abs().max())
    for k, g in tr.critic_grads().items():
        digest_check(g, gold, f"{state}/critic/grad/{k}", TOL_GRAD_BIAS if k.endswith(".bias") else TOL_GRAD,
                     f"critic grad {k}", abs_floor=1e-4, kinks=True)
    noise, mask_g, masks_d = PC.draw_generator_inputs(cfg, B)
    logs = tr.generator_update(real_bt, labels, noise, mask_g, masks_d, update=False)
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
    # drop-in forward API in eval mode (running statistics, no dropout)
    gen.eval()
    critic.eval()
    real = real_bt.reshape(B, T, cfg["output_size"]).permute(0, 2, 1).contiguous()
    noise = torch.randn(B, T, cfg["input_vector_size"], generator=torch.Generator().manual_seed(3))
    fake = gen(noise.to(dev), labels.to(dev)).cpu()
    ref = torch.from_numpy(gold[f"{state}/eval/fake"])
    assert fake.shape == ref.shape
    assert float((fake - ref).abs().max()) < TOL_FP32 * float(ref.abs().max())
    score = critic(real.to(dev), labels.to(dev)).cpu()
    ref = torch.from_numpy(gold[f"{state}/eval/score"])
    assert score.shape == ref.shape
    assert float((score - ref).abs().max()) < TOL_FP32 * max(1.0, float(ref.abs().max()))


@pytest.mark.gpu
def test_cuda_phase2_cond_batch24_vs_oracle():
    """default.yaml batch (24), all four labels present: one critic iteration with Adam, then a generator update,
    against the oracle."""
    from music2dance_b200.phase2_cond import Phase2CondTrainer
    dev = "cuda:0"
    cfg = PC.make_cfg()
    B, T = cfg["batch_size"], cfg["stick_length"]
    gen, critic = _build(cfg, "perturbed", dev)
    G = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    D = {k: v.detach().cpu().clone() for k, v in critic.state_dict().items()}
    tr = Phase2CondTrainer(gen, critic, cfg, B)
    real_bt = PC.synthetic_poses(B, T, 5)
    labels = PC.synthetic_labels(B, 11)
    assert len(set(labels.tolist())) == 4
    torch.manual_seed(9)
    ins = PC.draw_critic_inputs(cfg, B)
    o = PC.critic_iteration(G, D, cfg, real_bt, labels, *ins)
    logs = tr.critic_iteration(real_bt, labels, *ins, update=True)
    for k in ("loss_critic", "gp", "w_dist"):
        scalar_check(logs[k], o[k], 1e-3, k)
    for k, g in tr.critic_grads().items():
        ref = o["grads"][k]
        ref = torch.zeros_like(g) if ref is None else ref
        err = float((g - ref).norm()) / max(float(ref.norm()), 1e-6)
        assert err < (5e-2 if k.endswith(".bias") else 2e-2), (k, err)
    st = O.AdamState.__new__(O.AdamState)
    st.lr, st.b1, st.b2, st.eps, st.t = cfg["lr_critic"], 0.9, 0.999, 1e-8, {}
    st.m = {k: torch.zeros_like(D[k]) for k in PC.trainable(D)}
    st.v = {k: torch.zeros_like(D[k]) for k in PC.trainable(D)}
    with torch.no_grad():
        st.step(D, o["grads"])
    ins = PC.draw_generator_inputs(cfg, B)
    o = PC.generator_update(G, D, cfg, real_bt, labels, *ins)
    logs = tr.generator_update(real_bt, labels, *ins, update=True)
    scalar_check(logs["loss_gen"], o["loss_gen"], 2e-2, "loss_gen after one critic Adam step")
    scalar_check(logs["tv"], o["tv"], 1e-3, "tv")
    g, ref = tr.generator_grads()["embed_label.weight"], o["grads"]["embed_label.weight"]
    assert float((g - ref).norm()) < 5e-2 * float(ref.norm()), "generator embedding gradient"


@pytest.mark.gpu
def test_cuda_embed_kernels_vs_indexing():
    """m2d_embed_rows / m2d_embed_grad against torch indexing on the host (bit-exact lookup; the gradient is an fp64
    tree sum rounded once, compared at 1e-6), plus the out-of-range label flag."""
    from music2dance_b200 import ops
    from music2dance_b200.ops import Mat
    dev = "cuda:0"
    g = torch.Generator().manual_seed(2)
    B, T, E, W = 9, 17, 4, 11                                   # W = width of the concatenated operand
    table = torch.randn(4, E, generator=g)
    labels = torch.randint(0, 4, (B,), generator=g)
    y = torch.full((B * T, W), -7.0, device=dev)
    ym = Mat.of(y.view(-1), 1, B * T, W)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.embed_rows(table.to(dev), labels.to(dev), ym.cols_slice(W - E, W), B, T, 4, err)
    want = table[labels].unsqueeze(1).expand(-1, T, -1).reshape(B * T, E)
    assert torch.equal(y[:, W - E:].cpu(), want) and int(err.item()) == 0
    assert float((y[:, :W - E] + 7.0).abs().max()) == 0.0        # the other columns are untouched
    dy = torch.randn(B * T, W, generator=g)
    dt = torch.full((4, E), 3.0, device=dev)
    ops.embed_grad(Mat.of(dy.to(dev).view(-1), 1, B * T, W).cols_slice(W - E, W), labels.to(dev), dt, B, T, 4)
    ref = torch.zeros(4, E, dtype=torch.float64)
    ref.index_add_(0, labels, dy[:, W - E:].double().view(B, T, E).sum(1))
    assert float((dt.cpu().double() - ref).abs().max()) < 1e-6 * max(1.0, float(ref.abs().max()))
    ops.embed_grad(Mat.of(dy.to(dev).view(-1), 1, B * T, W).cols_slice(W - E, W), labels.to(dev), dt, B, T, 4,
                   scale=0.5, beta=1.0)
    assert float((dt.cpu().double() - 1.5 * ref).abs().max()) < 2e-6 * max(1.0, float(ref.abs().max()))
    bad = labels.clone()
    bad[3] = 4
    ops.embed_rows(table.to(dev), bad.to(dev), ym.cols_slice(W - E, W), B, T, 4, err)
    assert int(err.item()) == 1


@pytest.mark.gpu
def test_cuda_phase2_cond_rejects_bad_label():
    dev = "cuda:0"
    cfg = PC.make_cfg()
    gen, critic = _build(cfg, "init", dev)
    gen.eval()
    with pytest.raises(IndexError):
        gen(torch.zeros(2, 120, cfg["input_vector_size"], device=dev), torch.tensor([0, 4], device=dev))
