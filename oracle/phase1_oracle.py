"""TEST INFRASTRUCTURE — CPU restatement of the reference's phase1 stick-figure WGAN-GP step (BASELINE.json
configs[0]; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package).

Follows phase1/archis/residual.py:4-71 (Generator, Discriminator, LinearBlock incl. its dead fc1/bn1 branch),
losses.py:5-54 (gradient_penalty, is_seq=False branch) and the loop body phase1/train_wgan-gp.py:77-109.
Functional, torch fp32, autograd for the gradients.  Pinned against the reference's own modules by
tests/golden/make_golden_phase1.py (fixture tests/golden/phase1.npz): with the same torch.manual_seed the
random draws below (noise, dropout masks, alpha) consume the CPU generator exactly like the reference."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F


def make_cfg(**over):
    """phase1/configs/b2l50s256.yaml."""
    cfg = dict(batch_size=64, gamma=10.0, nblocks_gen=2, nblocks_critic=2, latent_vector_size=50, size=256,
               output_size=69, lr_gen=1e-4, lr_critic=1e-4, n_critic_steps=5)
    cfg.update(over)
    return cfg


def _linear_init(out_f, in_f):
    """nn.Linear.reset_parameters: kaiming_uniform_(a=sqrt(5)) on the weight, then the bias."""
    w = torch.empty(out_f, in_f)
    torch.nn.init.kaiming_uniform_(w, a=5 ** 0.5)
    bound = 1 / in_f ** 0.5
    b = torch.empty(out_f).uniform_(-bound, bound)
    return w, b


def _bn_init(P, name, n):
    P[name + ".weight"], P[name + ".bias"] = torch.ones(n), torch.zeros(n)
    P[name + ".running_mean"], P[name + ".running_var"] = torch.zeros(n), torch.ones(n)
    P[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)


def init_generator_params(cfg):
    """Creation order of residual.py:5-18 (fc1, bn1, blocks[fc1, fc2, bn1, bn2], lastfc)."""
    P, S = OrderedDict(), cfg["size"]
    P["fc1.weight"], P["fc1.bias"] = _linear_init(S, cfg["latent_vector_size"])
    _bn_init(P, "bn1", S)
    for i in range(cfg["nblocks_gen"]):
        P[f"blocks.{i}.fc1.weight"], P[f"blocks.{i}.fc1.bias"] = _linear_init(S, S)
        P[f"blocks.{i}.fc2.weight"], P[f"blocks.{i}.fc2.bias"] = _linear_init(S, S)
        _bn_init(P, f"blocks.{i}.bn1", S)
        _bn_init(P, f"blocks.{i}.bn2", S)
    P["lastfc.weight"], P["lastfc.bias"] = _linear_init(cfg["output_size"], S)
    return P


def init_critic_params(cfg):
    P, S = OrderedDict(), cfg["size"]
    P["fc1.weight"], P["fc1.bias"] = _linear_init(S, cfg["output_size"])
    for i in range(cfg["nblocks_critic"]):
        P[f"blocks.{i}.fc1.weight"], P[f"blocks.{i}.fc1.bias"] = _linear_init(S, S)
        P[f"blocks.{i}.fc2.weight"], P[f"blocks.{i}.fc2.bias"] = _linear_init(S, S)
    P["lastfc.weight"], P["lastfc.bias"] = _linear_init(1, S)
    return P


def dropout_mask(shape):
    """at::dropout on CPU: noise = empty_like(x).bernoulli_(1 - p) (p = 0.5), then scaled by 1/(1-p)."""
    return torch.empty(shape).bernoulli_(0.5)


def _bn(P, name, x, train):
    if train:
        P[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, P[name + ".running_mean"], P[name + ".running_var"], P[name + ".weight"],
                        P[name + ".bias"], train, 0.1, 1e-5)


def generator_forward(P, cfg, noise, mask, train=True):
    """residual.py:20-24; `mask` (B, size) 0/1 dropout mask (None in eval mode)."""
    x = F.relu(_bn(P, "bn1", F.linear(noise, P["fc1.weight"], P["fc1.bias"]), train))
    for i in range(cfg["nblocks_gen"]):
        b = f"blocks.{i}"
        _bn(P, b + ".bn1", F.linear(x, P[b + ".fc1.weight"], P[b + ".fc1.bias"]), train)      # dead branch (Q1)
        x = x + F.relu(_bn(P, b + ".bn2", F.linear(x, P[b + ".fc2.weight"], P[b + ".fc2.bias"]), train))
    if mask is not None:
        x = x * (mask * 2.0)
    return F.linear(x, P["lastfc.weight"], P["lastfc.bias"])


def critic_forward(P, cfg, x, mask):
    """residual.py:41-45; x (B, 23, 3) or (B, 69)."""
    x = F.relu(F.linear(x.reshape(x.shape[0], -1), P["fc1.weight"], P["fc1.bias"]))
    for i in range(cfg["nblocks_critic"]):
        b = f"blocks.{i}"
        x = x + F.relu(F.linear(x, P[b + ".fc2.weight"], P[b + ".fc2.bias"]))                # fc1 branch is dead
    if mask is not None:
        x = x * (mask * 2.0)
    return F.linear(x, P["lastfc.weight"], P["lastfc.bias"])


def gradient_penalty(P, cfg, real, fake, alpha, mask):
    """losses.py:13-54, is_seq=False / audio=None / lp=False."""
    B = real.shape[0]
    a = alpha.view(B, 1)
    x = (a * real.reshape(B, -1).detach() + (1 - a) * fake.reshape(B, -1).detach()).view(B, 23, 3)
    x.requires_grad_(True)
    out = critic_forward(P, cfg, x, mask)
    g = torch.autograd.grad(out, x, torch.ones_like(out), create_graph=True)[0].reshape(B, -1)
    return ((torch.sqrt((g ** 2).sum(1) + 1e-12) - 1) ** 2).mean()


def draw_critic_randoms(cfg, B):
    """RNG order of train_wgan-gp.py:84-89 on the CPU generator."""
    S = cfg["size"]
    noise = torch.randn(B, cfg["latent_vector_size"])
    mask_g = dropout_mask((B, S))
    alpha = torch.rand(B, 1)
    masks_d = [dropout_mask((B, S)) for _ in range(3)]          # interpolates, real, fake
    return noise, mask_g, alpha, masks_d


def draw_gen_randoms(cfg, B):
    """train_wgan-gp.py:99-102."""
    S = cfg["size"]
    noise = torch.randn(B, cfg["latent_vector_size"])
    mask_g = dropout_mask((B, S))
    masks_d = [dropout_mask((B, S)) for _ in range(2)]          # real, fake
    return noise, mask_g, masks_d


def _leaf(P):
    return OrderedDict((k, v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                       for k, v in P.items())


def trainable(P):
    return [k for k, v in P.items() if v.is_floating_point() and "running" not in k]


def critic_iteration(G, D, cfg, real, noise, mask_g, alpha, masks_d):
    """train_wgan-gp.py:80-94 (without the optimiser step).  Returns scalars + critic gradients."""
    with torch.no_grad():
        fake = generator_forward(G, cfg, noise, mask_g, train=True)
    Dl = _leaf(D)
    gp = gradient_penalty(Dl, cfg, real, fake, alpha, masks_d[0])
    err_real = critic_forward(Dl, cfg, real, masks_d[1]).mean()
    err_fake = critic_forward(Dl, cfg, fake.detach(), masks_d[2]).mean()
    err = err_fake - err_real + cfg["gamma"] * gp
    names = trainable(D)
    gl = torch.autograd.grad(err, [Dl[k] for k in names], allow_unused=True)
    return dict(loss_critic=float(err), gp=float(gp), w_dist=float(err_fake - err_real), fake=fake,
                grads=OrderedDict(zip(names, gl)))


def generator_update(G, D, cfg, real, noise, mask_g, masks_d):
    """train_wgan-gp.py:97-106 (without the optimiser step)."""
    Gl = _leaf(G)
    fake = generator_forward(Gl, cfg, noise, mask_g, train=True)
    for k in G:                                                   # running statistics advanced in Gl's shared buffers
        if "running" in k or "num_batches" in k:
            G[k] = Gl[k]
    err_real = critic_forward(D, cfg, real, masks_d[0]).mean()
    err_fake = critic_forward(D, cfg, fake, masks_d[1]).mean()
    err = err_real - err_fake
    names = trainable(G)
    gl = torch.autograd.grad(err, [Gl[k] for k in names], allow_unused=True)
    return dict(loss_gen=float(err), fake=fake.detach(), grads=OrderedDict(zip(names, gl)))


def synthetic_poses(B, seed):
    """MinMax-scaled poses live in [0, 1] (utils.py:26-31)."""
    return torch.rand(B, 23, 3, generator=torch.Generator().manual_seed(seed))
