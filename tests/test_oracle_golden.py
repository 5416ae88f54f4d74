"""CPU: the oracle restatement against the golden fixtures generated from the
unmodified reference modules (tests/golden/make_golden.py)."""
import pytest
import torch

from oracle import phase3_oracle as O
from tests.parity import (ALPHA_SEED, B_GOLD, DATA_SEED, TOL_FP32, TOL_GEN_GRAD_E2E, TOL_GRAD, VARIANTS, digest_check, load_golden,
                          scalar_check)


def _state(cfg, state):
    torch.manual_seed(0)
    G, D = O.init_generator_params(cfg), O.init_critic_params(cfg)
    if state == "perturbed":
        O.perturb_params(G)
        O.perturb_params(D)
    return G, D


@pytest.mark.parametrize("state", ["init", "perturbed"])
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_oracle_matches_reference_fixtures(variant, state):
    torch.set_num_threads(8)
    cfg = O.make_cfg(**VARIANTS[variant])
    gold = load_golden(variant)
    G, D = _state(cfg, state)
    for k, v in list(G.items()) + list(D.items()):
        digest_check(v, gold, f"{state}/init/{k}", 1e-7, f"init {k}")
    real, audio, noise, _, noise_g = O.synthetic_batch(cfg, B_GOLD, int(gold[f"{state}/data_seed"]))
    torch.manual_seed(ALPHA_SEED)
    alpha = torch.rand(B_GOLD, 1)
    o = O.critic_iteration(G, D, cfg, real, audio, noise, alpha, None)
    for k in ("loss_critic", "gp", "w_dist", "err_real", "err_fake"):
        scalar_check(o[k], gold[f"{state}/critic/{k}"], TOL_FP32, k)
    ref_fake = torch.from_numpy(gold[f"{state}/critic/fake"])
    assert float((o["fake"] - ref_fake).abs().max()) < TOL_FP32 * float(ref_fake.abs().max())
    for k, g in o["grads"].items():
        digest_check(g, gold, f"{state}/critic/grad/{k}", TOL_GRAD, f"critic grad {k}", abs_floor=1e-4, kinks=True)
    for k, v in G.items():
        if "running" in k or "num_batches" in k:
            digest_check(v, gold, f"{state}/critic/genbuf/{k}", TOL_FP32, f"bn buffer {k}")
    o = O.generator_update(G, D, cfg, real, audio, noise_g, None)
    for k in ("loss_gen", "l1", "tv", "err_real", "err_fake"):
        scalar_check(o[k], gold[f"{state}/gen/{k}"], TOL_FP32, k)
    skip = set(O.pre_bn_bias_names(G))
    for k, g in o["grads"].items():
        if g is None:
            assert f"{state}/gen/nograd/{k}" in gold.files, k
        elif k not in skip:
            digest_check(g, gold, f"{state}/gen/grad/{k}", TOL_GEN_GRAD_E2E, f"gen grad {k}", abs_floor=1e-4, kinks=True)


def test_windowing_bit_exact_vs_fixture():
    gold = load_golden("default")
    g = torch.Generator().manual_seed(5)
    for name in "abcd":
        n, win, stride = [int(v) for v in gold[f"window/{name}/args"]]
        x = torch.rand(2, n, generator=g)
        s = O.slice_audio_batch(x, win, stride, win - stride)
        assert list(s.shape) == [int(v) for v in gold[f"window/{name}/shape"]]
        d = O.tensor_digest(s, 256)
        assert d["sum"] == float(gold[f"window/{name}/sum"])
        assert torch.equal(d["samples"], torch.from_numpy(gold[f"window/{name}/samples"]))
        assert torch.equal(O.slice_audio_batch(x[0], win, stride, win - stride), s[0])


def test_edge_windows():
    # audio shorter than / equal to one window, ragged tail
    for n, win, stride in [(3200, 3200, 640), (640, 3200, 640), (3201, 3200, 640), (2, 4, 2)]:
        x = torch.arange(1, n + 1, dtype=torch.float32)
        s = O.slice_audio_batch(x, win, stride, win - stride)
        pad = win - stride
        assert s.shape[0] == (n + pad - win) // stride + 1
        full = torch.cat([torch.zeros(pad // 2), x, torch.zeros(pad - pad // 2)])
        for i in range(s.shape[0]):
            assert torch.equal(s[i], full[i * stride:i * stride + win])
