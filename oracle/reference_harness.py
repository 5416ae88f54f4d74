"""Drive the *unmodified* reference modules (read-only at /root/reference) through
one phase3 train-step body, for validating the oracle restatement and for
generating golden fixtures.  Only usable where /root/reference exists (the build
container); nothing that runs on the GPU box imports this file.

The reference script phase3/train.py cannot run as-is (dataset absent,
``yaml.load`` without Loader), so the loop body train.py:186-237 is re-stated
here around the reference's own ``archis.default`` / ``losses`` / ``utils``.
``librosa`` (absent, only used by the dataset loader utils.py:185) is stubbed.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF = os.environ.get("M2D_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "phase3", "archis"))


def import_reference():
    if "librosa" not in sys.modules:
        sys.modules["librosa"] = types.ModuleType("librosa")
    for p in (REF, os.path.join(REF, "phase3")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import archis.default as archis      # noqa: E402
    import losses                        # noqa: E402
    import utils                         # noqa: E402
    return archis, losses, utils


def build_models(cfg, seed=0):
    """train.py:35,87-98: manual_seed(0); generator first, then critic."""
    archis, _, _ = import_reference()
    torch.manual_seed(seed)
    gen = archis.SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"],
                                   cfg["latent_vector_size"], cfg["size"], cfg["output_size"],
                                   cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"],
                                   cfg["enc_type"], cfg["activ"], "cpu")
    cls = archis.AblatedSequenceDiscriminator if cfg["ablated"] else archis.SequenceDiscriminator
    critic = cls(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                 init_ker=cfg["init_kernel"], activ=cfg["activ"], device="cpu")
    return gen, critic


def critic_iteration(gen, critic, cfg, real_bt, audio, noise, alpha_seed, optim_critic=None):
    """train.py:187-216 with the generator noise passed explicitly and the GP
    alpha drawn by the reference itself right after ``manual_seed(alpha_seed)``."""
    _, losses, utils = import_reference()
    B, T, O = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    gen.train()
    critic.zero_grad()
    slices = utils.slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"],
                                     cfg["pad_samples"])
    aud = audio.clone().unsqueeze(1)
    fake = gen(slices, [T] * B, noise=noise)
    fake = fake.view(B, T, O).permute(0, 2, 1).contiguous()
    real = real_bt.view(B, T, O).permute(0, 2, 1).contiguous()
    torch.manual_seed(alpha_seed)
    if cfg["ablated"]:
        gp = losses.gradient_penalty(critic, B, real, fake, is_seq=True, lp=False, device="cpu")
        err_real = torch.mean(critic(real))
        err_fake = torch.mean(critic(fake.detach()))
    else:
        gp = losses.gradient_penalty(critic, B, real, fake, aud, is_seq=True, lp=False, device="cpu")
        err_real = torch.mean(critic(real, aud))
        err_fake = torch.mean(critic(fake.detach(), aud))
    err = err_fake - err_real + cfg["gamma"] * gp
    err.backward(retain_graph=True)
    grads = {k: (None if p.grad is None else p.grad.detach().clone())
             for k, p in critic.named_parameters()}
    if optim_critic is not None:
        optim_critic.step()
    return dict(loss_critic=err.item(), gp=gp.item(), w_dist=(err_fake - err_real).item(),
                err_real=err_real.item(), err_fake=err_fake.item(), fake=fake.detach(),
                grads=grads, slices=slices)


def generator_update(gen, critic, cfg, real_bt, audio, noise, optim_gen=None):
    """train.py:222-237."""
    _, losses, utils = import_reference()
    B, T, O = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    gen.train()
    gen.zero_grad()
    slices = utils.slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"],
                                     cfg["pad_samples"])
    aud = audio.clone().unsqueeze(1)
    real = real_bt.view(B, T, O).permute(0, 2, 1).contiguous()
    fake = gen(slices, [T] * B, noise=noise).view(B, T, O).permute(0, 2, 1)
    l1 = torch.nn.L1Loss(reduction="mean")(real, fake)
    if cfg["ablated"]:
        err_real, err_fake = torch.mean(critic(real)), torch.mean(critic(fake))
    else:
        err_real, err_fake = torch.mean(critic(real, aud)), torch.mean(critic(fake, aud))
    tv = losses.tv_loss(fake)
    err = err_real - err_fake + cfg["beta"] * l1 + cfg["eta"] * tv
    err.backward()
    grads = {k: (None if p.grad is None else p.grad.detach().clone())
             for k, p in gen.named_parameters()}
    if optim_gen is not None:
        optim_gen.step()
    return dict(loss_gen=err.item(), l1=l1.item(), tv=tv.item(), err_real=err_real.item(),
                err_fake=err_fake.item(), fake=fake.detach(), grads=grads)
