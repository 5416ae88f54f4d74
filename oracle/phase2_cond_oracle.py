"""TEST INFRASTRUCTURE — CPU restatement of the reference's phase2 CONDITIONAL (dance-type label) sequence WGAN
modules and step (BASELINE.json configs[2]; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this).

Follows phase2/archis/conditional.py:6-49 (SequenceGenerator: Embedding(4,4) label code concatenated to the noise at
every frame -> GRU -> FrameDecoder with Dropout(0.5) before ``lastfc`` :110-131; SequenceDiscriminator: the same
label code concatenated as 4 extra input channels -> conv1 -> TemporalBlocks -> Dropout(0.5) -> lastconv) and
losses.py:13-50 (gradient_penalty, is_seq=True, lp=True).

The reference's driver for these modules, phase2/train_conditional.py, is bit-rotted against them (it constructs the
critic with other arguments and expects an auxiliary-classifier critic returning a tuple, :76-77,130-137), so the
step restated here is the loop body of the sibling phase2/train.py:134-171 with the labels threaded through both
networks — ``critic(x, labels)`` and ``gen(noise, labels)`` with the real batch's labels — and the penalty taken on
``lambda x: critic(x, labels)`` (lp=True as train_conditional.py:127-128 asks).  Dropout masks are explicit inputs:
at::dropout on CPU draws ``empty_like(x).bernoulli_(0.5)`` from the global generator, reproduced by ``dropout_mask``
in the reference's call order (``draw_*``).  Pinned against the reference's modules by
tests/golden/make_golden_phase2_cond.py."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import phase3_oracle as O

N_CLASSES, EMBED = 4, 4                       # conditional.py:13,31: nn.Embedding(4, 4)


def make_cfg(**over):
    """phase2/configs/default.yaml (the only phase2 config the reference ships)."""
    cfg = dict(batch_size=24, stick_length=120, gamma=10.0, eta=50.0, nblocks_gen=2, nblocks_critic=3,
               input_vector_size=50, latent_vector_size=50, n_cells=3, size=256, channels=128, output_size=69,
               lr_gen=5e-4, lr_critic=5e-4, n_critic_steps=8, init_kernel=25)
    cfg.update(over)
    return cfg


def _embedding():
    """nn.Embedding's default init (normal_(0, 1)) is the FIRST draw of both constructors (conditional.py:13,31);
    utils.initialize_weights (utils.py:267-313) leaves Embedding untouched."""
    return nn.Embedding(N_CLASSES, EMBED).weight.detach().clone()


def init_generator_params(cfg):
    """conditional.py:7-16: embed_label, GRU(input+4 -> latent), FrameDecoder, then xavier_normal_ (O._materialise)."""
    S = cfg["size"]
    P = OrderedDict([("embed_label.weight", _embedding())])
    specs = [("noise_gen.rnn", "gru", (cfg["input_vector_size"] + EMBED, cfg["latent_vector_size"], cfg["n_cells"])),
             ("decoder.fc1", "lin", (cfg["latent_vector_size"], S)), ("decoder.bn1", "bn", (S,))]
    for b in range(cfg["nblocks_gen"]):
        q = f"decoder.blocks.{b}."
        specs += [(q + "fc1", "lin", (S, S)), (q + "fc2", "lin", (S, S)), (q + "bn1", "bn", (S,)), (q + "bn2", "bn", (S,))]
    specs.append(("decoder.lastfc", "lin", (S, cfg["output_size"])))
    P.update(O._materialise(specs))
    return P


def init_critic_params(cfg):
    """conditional.py:29-42."""
    Ci, Ch, T, k0 = cfg["output_size"] + EMBED, cfg["channels"], cfg["stick_length"], cfg["init_kernel"]
    P = OrderedDict([("embed_label.weight", _embedding())])
    specs = [("conv1", "conv", (Ci, Ch, k0, 1, int((k0 - 1) / 2)))]
    for b in range(cfg["nblocks_critic"]):
        specs += [(f"blocks.{b}.conv1", "conv", (Ch, Ch, 7, 1, 3)), (f"blocks.{b}.conv2", "conv", (Ch, Ch, 7, 1, 3))]
    specs.append(("lastconv", "conv", (Ch, 1, T, 1, 0)))
    P.update(O._materialise(specs))
    return P


def dropout_mask(shape):
    """at::dropout (CPU, train): noise = empty_like(x).bernoulli_(1 - p), output = x * noise / (1 - p); p = 0.5."""
    return torch.empty(shape).bernoulli_(0.5)


def generator_forward(P, cfg, noise, labels, mask, train=True):
    """conditional.py:18-25: noise (B, T, input), labels (B,) int64, mask (B*T, size) 0/1 or None -> (B*T, 69)."""
    B, T, _ = noise.shape
    y = P["embed_label.weight"][labels].unsqueeze(1).expand(-1, T, -1)
    x = O.gru_forward(P, "noise_gen.rnn", torch.cat((noise, y), 2), cfg["n_cells"])
    x = x.reshape(B * T, cfg["latent_vector_size"])
    # FrameDecoder (conditional.py:124-128) = the phase3 decoder with dropout in front of lastfc
    x = O._relu(O._bn(P, "decoder.bn1", F.linear(x, P["decoder.fc1.weight"], P["decoder.fc1.bias"]), train))
    for b in range(cfg["nblocks_gen"]):
        q = f"decoder.blocks.{b}."
        O._bn(P, q + "bn1", F.linear(x, P[q + "fc1.weight"], P[q + "fc1.bias"]), train)          # dead branch (Q1)
        x = x + O._relu(O._bn(P, q + "bn2", F.linear(x, P[q + "fc2.weight"], P[q + "fc2.bias"]), train))
    if mask is not None:
        x = x * (mask * 2.0)
    return F.linear(x, P["decoder.lastfc.weight"], P["decoder.lastfc.bias"])


def critic_forward(P, cfg, x, labels, mask):
    """conditional.py:44-49: x (B, 69, T), labels (B,), mask (B, channels, T) 0/1 or None -> (B, 1)."""
    k0 = P["conv1.weight"].shape[-1]
    y = P["embed_label.weight"][labels].unsqueeze(-1).expand(-1, -1, x.size(-1))
    x = F.relu(F.conv1d(torch.cat((x, y), 1), P["conv1.weight"], P["conv1.bias"], padding=(k0 - 1) // 2))
    for b in range(cfg["nblocks_critic"]):
        h = F.relu(F.conv1d(x, P[f"blocks.{b}.conv1.weight"], P[f"blocks.{b}.conv1.bias"], padding=3))
        h = F.relu(F.conv1d(h, P[f"blocks.{b}.conv2.weight"], P[f"blocks.{b}.conv2.bias"], padding=3))
        x = x + h
    if mask is not None:
        x = x * (mask * 2.0)
    return F.conv1d(x, P["lastconv.weight"], P["lastconv.bias"]).squeeze(1)


def gradient_penalty_lp(P, cfg, real, fake, labels, alpha, mask):
    """losses.py:13-50 with is_seq=True, lp=True on ``lambda x: critic(x, labels)``."""
    B = real.shape[0]
    a = alpha.view(B, 1)
    x = (a * real.reshape(B, -1).detach() + (1 - a) * fake.reshape(B, -1).detach()).view(B, 69, -1)
    x.requires_grad_(True)
    out = critic_forward(P, cfg, x, labels, mask)
    g = torch.autograd.grad(out, x, torch.ones_like(out), create_graph=True)[0].reshape(B, -1)
    bgrad = g.norm(2, dim=1) - 1
    bgrad = torch.where(bgrad < 0, torch.zeros_like(bgrad), bgrad)
    return (bgrad ** 2).mean()


def draw_critic_inputs(cfg, B):
    """Random draws of one critic iteration in the reference's order: noise, the generator's decoder dropout,
    alpha (losses.py:15), then the critic's dropout for the interpolates, the real and the fake evaluation."""
    T, S, Ch = cfg["stick_length"], cfg["size"], cfg["channels"]
    noise = torch.randn(B, T, cfg["input_vector_size"])
    mask_g = dropout_mask((B * T, S))
    alpha = torch.rand(B, 1)
    masks_d = [dropout_mask((B, Ch, T)) for _ in range(3)]
    return noise, mask_g, alpha, masks_d


def draw_generator_inputs(cfg, B):
    """noise, decoder dropout, then the critic's dropout for the real and the fake evaluation."""
    T, S, Ch = cfg["stick_length"], cfg["size"], cfg["channels"]
    noise = torch.randn(B, T, cfg["input_vector_size"])
    mask_g = dropout_mask((B * T, S))
    masks_d = [dropout_mask((B, Ch, T)) for _ in range(2)]
    return noise, mask_g, masks_d


def _leaf(P):
    return OrderedDict((k, v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                       for k, v in P.items())


def trainable(P):
    return [k for k, v in P.items() if v.is_floating_point() and "running" not in k]


def critic_iteration(G, D, cfg, real_bt, labels, noise, mask_g, alpha, masks_d):
    """phase2/train.py:134-153 with labels (no optimiser step).  real_bt (B, T, 23, 3)."""
    B, T, Oo = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    with torch.no_grad():
        fake = generator_forward(G, cfg, noise, labels, mask_g, train=True).view(B, T, Oo).permute(0, 2, 1).contiguous()
    real = real_bt.reshape(B, T, Oo).permute(0, 2, 1).contiguous()
    Dl = _leaf(D)
    gp = gradient_penalty_lp(Dl, cfg, real, fake, labels, alpha, masks_d[0])
    err_real = critic_forward(Dl, cfg, real, labels, masks_d[1]).mean()
    err_fake = critic_forward(Dl, cfg, fake.detach(), labels, masks_d[2]).mean()
    err = err_fake - err_real + cfg["gamma"] * gp
    names = trainable(D)
    gl = torch.autograd.grad(err, [Dl[k] for k in names], allow_unused=True)
    return dict(loss_critic=float(err.detach()), gp=float(gp.detach()), w_dist=float((err_fake - err_real).detach()),
                fake=fake, grads=OrderedDict(zip(names, gl)))


def generator_update(G, D, cfg, real_bt, labels, noise, mask_g, masks_d):
    """phase2/train.py:159-171 with labels (no optimiser step)."""
    B, T, Oo = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    real = real_bt.reshape(B, T, Oo).permute(0, 2, 1).contiguous()
    Gl = _leaf(G)
    fake = generator_forward(Gl, cfg, noise, labels, mask_g, train=True).view(B, T, Oo).permute(0, 2, 1)
    for k in G:
        if "running" in k or "num_batches" in k:
            G[k] = Gl[k]
    err_real = critic_forward(D, cfg, real, labels, masks_d[0]).mean()
    err_fake = critic_forward(D, cfg, fake, labels, masks_d[1]).mean()
    tv = O.tv_loss(fake)
    err = err_real - err_fake + cfg["eta"] * tv
    names = trainable(G)
    gl = torch.autograd.grad(err, [Gl[k] for k in names], allow_unused=True)
    return dict(loss_gen=float(err.detach()), tv=float(tv.detach()), fake=fake.detach(),
                grads=OrderedDict(zip(names, gl)))


def synthetic_poses(B, T, seed):
    return torch.rand(B, T, 23, 3, generator=torch.Generator().manual_seed(seed))


def synthetic_labels(B, seed):
    """Dance-type labels 0..3 (utils.SequenceDataset returns the class index of the sequence's dance type)."""
    return torch.randint(0, N_CLASSES, (B,), generator=torch.Generator().manual_seed(seed))
