"""Validation / evaluation path: phase3/train.py:245-261 (eval-mode generator, mean L1) and
losses.jerkiness (phase3/test.py:85-100).  Oracle against values produced by the reference's own modules
(tests/golden/phase3_eval.npz); GPU: Phase3Trainer.validate and the drop-in jerkiness against both.
Tolerance: fp32 forward, TOL_FP32 = 2e-4 relative (the north star allows 1e-3)."""
import os

import numpy as np
import pytest
import torch

from oracle import phase3_oracle as O
from tests.parity import TOL_FP32, scalar_check

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phase3_eval.npz")
ENCODERS = ["default", "wavegan", "unet"]


def _state(cfg):
    torch.manual_seed(0)                                          # phase3/train.py:35
    G = O.init_generator_params(cfg)
    O.perturb_params(G)
    return G


@pytest.mark.parametrize("enc", ENCODERS)
def test_oracle_validation_and_jerkiness_vs_reference(enc):
    gold = np.load(GOLD)
    cfg = O.make_cfg(enc_type=enc)
    B = int(gold["B"])
    real_bt, audio, noise, _, _ = O.synthetic_batch(cfg, B, int(gold["seed"]))
    l1, fake = O.validation_l1(_state(cfg), cfg, real_bt, audio, noise)
    ref_fake = torch.from_numpy(gold[f"{enc}/fake"])
    assert float((fake - ref_fake).abs().max()) < 1e-5 * float(ref_fake.abs().max())
    scalar_check(l1, gold[f"{enc}/l1_val"], 1e-6, "l1_val")
    scalar_check(O.jerkiness(ref_fake), gold[f"{enc}/jerk_fake"], 1e-6, "jerkiness(fake)")
    real = real_bt.reshape(B, cfg["stick_length"], 69).permute(0, 2, 1)
    scalar_check(O.jerkiness(real), gold[f"{enc}/jerk_real"], 1e-6, "jerkiness(real)")


@pytest.mark.gpu
@pytest.mark.parametrize("enc", ENCODERS)
def test_trainer_validate_and_jerkiness(enc):
    from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator
    from music2dance_b200.losses import jerkiness
    from music2dance_b200.trainer import Phase3Trainer
    dev = "cuda:0"
    gold = np.load(GOLD)
    cfg = O.make_cfg(enc_type=enc, n_critic_steps=1)
    torch.manual_seed(0)
    gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"],
                            cfg["size"], cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"],
                            cfg["enc_type"], cfg["activ"], dev)
    critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                                   init_ker=cfg["init_kernel"], activ=cfg["activ"], device=dev)
    sd = {k: v.cpu() for k, v in gen.state_dict().items()}
    O.perturb_params(sd)
    gen.load_state_dict(sd, strict=True)
    before = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    tr = Phase3Trainer(gen, critic, cfg, 2, use_graphs=False)          # training batch 2, validation batch 3
    B = int(gold["B"])
    real_bt, audio, noise, _, _ = O.synthetic_batch(cfg, B, int(gold["seed"]))
    l1, fake = tr.validate(real_bt, audio, noise)
    T, Oo = cfg["stick_length"], cfg["output_size"]
    fake = fake.view(B, T, Oo).permute(0, 2, 1)
    ref_fake = torch.from_numpy(gold[f"{enc}/fake"])
    assert float((fake.cpu() - ref_fake).abs().max()) < TOL_FP32 * float(ref_fake.abs().max())
    scalar_check(l1, gold[f"{enc}/l1_val"], TOL_FP32, "l1_val vs reference")
    scalar_check(jerkiness(fake), gold[f"{enc}/jerk_fake"], 1e-3, "jerkiness(fake) vs reference")
    scalar_check(jerkiness(ref_fake.to(dev)), gold[f"{enc}/jerk_fake"], 1e-5, "jerkiness kernel")
    real = real_bt.reshape(B, T, Oo).permute(0, 2, 1).to(dev)
    scalar_check(jerkiness(real), gold[f"{enc}/jerk_real"], 1e-5, "jerkiness(real)")
    flat = ref_fake.permute(0, 2, 1).reshape(1, -1, Oo).permute(0, 2, 1).to(dev)      # phase3/test.py:92 form
    scalar_check(jerkiness(flat), gold[f"{enc}/jerk_fake_flat"], 1e-5, "jerkiness(flat)")
    # eval mode: running statistics and parameters untouched
    for k, v in gen.state_dict().items():
        assert torch.equal(v.cpu(), before[k]), f"validate() modified {k}"
