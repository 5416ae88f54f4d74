"""TEST INFRASTRUCTURE — CPU restatement of the reference's phase2 unconditional sequence WGAN-LP step
(BASELINE.json configs[1]; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this).

Follows phase2/archis/default.py:5-49,88-163 (SequenceGenerator = GRU noise generator + FrameDecoder,
SequenceDiscriminator = conv1 + TemporalBlocks + lastconv), losses.py:13-50 (gradient_penalty, is_seq=True,
lp=True) and the loop body phase2/train.py:131-171.  Built from the phase3 oracle's pieces (same GRU, decoder,
convolution helpers).  Pinned against the reference's modules by tests/golden/make_golden_phase2.py."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import phase3_oracle as O


def make_cfg(**over):
    """phase2/configs/default.yaml."""
    cfg = dict(batch_size=24, stick_length=120, gamma=10.0, eta=50.0, nblocks_gen=2, nblocks_critic=3,
               input_vector_size=50, latent_vector_size=50, n_cells=3, size=256, channels=128, output_size=69,
               lr_gen=5e-4, lr_critic=5e-4, n_critic_steps=8, init_kernel=25)
    cfg.update(over)
    return cfg


def init_generator_params(cfg):
    """Constructor RNG order of SequenceGenerator (default.py:6-16): default inits in creation order (GRU, then the
    decoder's Linear layers), then initialize_weights' xavier_normal_ in module-traversal order — the phase3
    oracle's _materialise implements exactly that protocol."""
    S = cfg["size"]
    specs = [("noise_gen.rnn", "gru", (cfg["input_vector_size"], cfg["latent_vector_size"], cfg["n_cells"])),
             ("decoder.fc1", "lin", (cfg["latent_vector_size"], S)), ("decoder.bn1", "bn", (S,))]
    for b in range(cfg["nblocks_gen"]):
        q = f"decoder.blocks.{b}."
        specs += [(q + "fc1", "lin", (S, S)), (q + "fc2", "lin", (S, S)), (q + "bn1", "bn", (S,)), (q + "bn2", "bn", (S,))]
    specs.append(("decoder.lastfc", "lin", (S, cfg["output_size"])))
    return O._materialise(specs)


def init_critic_params(cfg):
    """SequenceDiscriminator (default.py:27-41)."""
    Ci, Ch, T, k0 = cfg["output_size"], cfg["channels"], cfg["stick_length"], cfg["init_kernel"]
    specs = [("conv1", "conv", (Ci, Ch, k0, 1, int((k0 - 1) / 2)))]
    for b in range(cfg["nblocks_critic"]):
        specs += [(f"blocks.{b}.conv1", "conv", (Ch, Ch, 7, 1, 3)), (f"blocks.{b}.conv2", "conv", (Ch, Ch, 7, 1, 3))]
    specs.append(("lastconv", "conv", (Ch, 1, T, 1, 0)))
    return O._materialise(specs)


def generator_forward(P, cfg, noise, train=True):
    """default.py:18-24: noise (B, T, input) -> (B*T, 69)."""
    B, T, _ = noise.shape
    x = O.gru_forward(P, "noise_gen.rnn", noise, cfg["n_cells"])
    return O.decoder_forward(P, x.reshape(B * T, cfg["latent_vector_size"]), cfg["nblocks_gen"], train)


def critic_forward(P, cfg, x):
    """default.py:43-49: x (B, 69, T) -> (B, 1)."""
    k0 = P["conv1.weight"].shape[-1]
    x = F.relu(F.conv1d(x, P["conv1.weight"], P["conv1.bias"], padding=(k0 - 1) // 2))
    for b in range(cfg["nblocks_critic"]):
        y = F.relu(F.conv1d(x, P[f"blocks.{b}.conv1.weight"], P[f"blocks.{b}.conv1.bias"], padding=3))
        y = F.relu(F.conv1d(y, P[f"blocks.{b}.conv2.weight"], P[f"blocks.{b}.conv2.bias"], padding=3))
        x = x + y
    return F.conv1d(x, P["lastconv.weight"], P["lastconv.bias"]).squeeze(1)


def gradient_penalty_lp(P, cfg, real, fake, alpha):
    """losses.py:13-50 with is_seq=True, lp=True."""
    B = real.shape[0]
    a = alpha.view(B, 1)
    x = (a * real.reshape(B, -1).detach() + (1 - a) * fake.reshape(B, -1).detach()).view(B, 69, -1)
    x.requires_grad_(True)
    out = critic_forward(P, cfg, x)
    g = torch.autograd.grad(out, x, torch.ones_like(out), create_graph=True)[0].reshape(B, -1)
    bgrad = g.norm(2, dim=1) - 1
    bgrad = torch.where(bgrad < 0, torch.zeros_like(bgrad), bgrad)
    return (bgrad ** 2).mean()


def _leaf(P):
    return OrderedDict((k, v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                       for k, v in P.items())


def trainable(P):
    return [k for k, v in P.items() if v.is_floating_point() and "running" not in k]


def critic_iteration(G, D, cfg, real_bt, noise, alpha):
    """phase2/train.py:134-153 (without the optimiser step).  real_bt (B, T, 23, 3)."""
    B, T, Oo = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    with torch.no_grad():
        fake = generator_forward(G, cfg, noise, train=True).view(B, T, Oo).permute(0, 2, 1).contiguous()
    real = real_bt.reshape(B, T, Oo).permute(0, 2, 1).contiguous()
    Dl = _leaf(D)
    gp = gradient_penalty_lp(Dl, cfg, real, fake, alpha)
    err_real = critic_forward(Dl, cfg, real).mean()
    err_fake = critic_forward(Dl, cfg, fake.detach()).mean()
    err = err_fake - err_real + cfg["gamma"] * gp
    names = trainable(D)
    gl = torch.autograd.grad(err, [Dl[k] for k in names], allow_unused=True)
    return dict(loss_critic=float(err), gp=float(gp), w_dist=float(err_fake - err_real), fake=fake,
                grads=OrderedDict(zip(names, gl)))


def generator_update(G, D, cfg, real_bt, noise):
    """phase2/train.py:159-171 (without the optimiser step)."""
    B, T, Oo = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    real = real_bt.reshape(B, T, Oo).permute(0, 2, 1).contiguous()
    Gl = _leaf(G)
    fake = generator_forward(Gl, cfg, noise, train=True).view(B, T, Oo).permute(0, 2, 1)
    for k in G:
        if "running" in k or "num_batches" in k:
            G[k] = Gl[k]
    err_real = critic_forward(D, cfg, real).mean()
    err_fake = critic_forward(D, cfg, fake).mean()
    tv = O.tv_loss(fake)
    err = err_real - err_fake + cfg["eta"] * tv
    names = trainable(G)
    gl = torch.autograd.grad(err, [Gl[k] for k in names], allow_unused=True)
    return dict(loss_gen=float(err), tv=float(tv), fake=fake.detach(), grads=OrderedDict(zip(names, gl)))


def synthetic_poses(B, T, seed):
    return torch.rand(B, T, 23, 3, generator=torch.Generator().manual_seed(seed))
