"""CPU oracle for the phase3 audio-to-dance WGAN-GP training step.

TEST INFRASTRUCTURE ONLY.  Nothing in ``music2dance_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may.  It is the checker, never the product path.

What it is: a *functional* restatement (plain torch fp32 on CPU, parameters in
an ordered dict keyed by the reference's ``state_dict`` names) of the reference
hot path:

    phase3/archis/default.py   generator / encoders / decoder / critic
    losses.py:5-60,76-82       gradient_penalty (seq branches), tv_loss
    utils.py:267-313,329-353   initialize_weights, slice_audio_batch
    phase3/train.py:186-237    critic iteration + generator update

The arithmetic of the reference lives in PyTorch (un-pinned; the installed
torch 2.11 is the operational pin), so the restatement uses
``torch.nn.functional`` primitives for conv / linear / batch-norm and writes
the GRU cell, WGAN-GP, losses and Adam out explicitly.

Parity pin: the reference ships no tests or golden vectors ("parity unpinned"
by the reference itself).  This oracle is pinned instead against the
*reference's own modules executed in the build container*:
``oracle/validate_vs_reference.py`` runs both on identical seeds, and
``tests/golden/make_golden.py`` stores reference outputs as fixtures that
``tests/test_oracle_golden.py`` re-checks wherever the reference is absent.
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------
# configuration (phase3/configs/default.yaml:1-37 + derived constants,
# phase3/train.py:45-83, utils.py:50-58)
# --------------------------------------------------------------------------

DEFAULT_CFG = dict(
    batch_size=7, window_size=0.2, seq_length=4.8, gamma=10.0, beta=1.0, eta=0.0,
    nblocks_gen=2, input_vector_size=250, latent_vector_size=250, n_cells=3,
    size=256, channels=128, output_size=69, lr_gen=2e-4, lr_critic=2e-4,
    n_critic_steps=8, code_size=100, noise_size=10, init_kernel=25,
    enc_type="default", ablated=False, activ="id",
    audio_rate=16000, video_rate=25,
)


def make_cfg(**over):
    cfg = dict(DEFAULT_CFG)
    cfg.update(over)
    cfg["stick_length"] = int(cfg["seq_length"] * cfg["video_rate"])          # utils.py:55
    cfg["audio_length"] = int(cfg["seq_length"] * cfg["audio_rate"])          # utils.py:56
    cfg["cutting_stride"] = int(cfg["audio_rate"] / cfg["video_rate"])        # utils.py:57
    cfg["audio_feat_samples"] = int(cfg["window_size"] * cfg["audio_rate"])   # train.py:82
    cfg["pad_samples"] = cfg["audio_feat_samples"] - cfg["cutting_stride"]    # train.py:83
    return cfg


# --------------------------------------------------------------------------
# parameter construction: same RNG consumption as the reference constructors
# (default torch init at creation, then xavier_normal_ over every Conv1d /
# Linear weight and GRU weight_* in module-traversal order, utils.py:267-313)
# --------------------------------------------------------------------------

def _gen_layer_specs(cfg):
    f, enc, out = 32, cfg["enc_type"], cfg["input_vector_size"]
    specs = []
    p = "audio_enc.model."
    if enc == "default":                                   # default.py:59-76
        chans = [1, f, 2 * f, 4 * f, 8 * f, 16 * f, 32 * f]
        specs.append((p + "conv_layers.0", "conv", (1, f, 250, 50, 124)))
        for i in range(1, 6):
            specs.append((p + f"conv_layers.{i}", "conv", (chans[i], chans[i + 1], 4, 2, 1)))
        specs.append((p + "conv_layers.6", "conv", (chans[6], out, 2, 1, 0)))
        for i in range(6):
            specs.append((p + f"activations.{i}.0", "bn", (chans[i + 1],)))
    elif enc == "wavegan":                                 # default.py:114-135
        chans = [1, f, 2 * f, 4 * f, 8 * f]
        for i in range(1, 5):
            specs.append((p + f"l{i}", "conv", (chans[i - 1], chans[i], 25, 4, 0)))
            specs.append((p + f"bn{i}", "bn", (chans[i],)))
        specs.append((p + "l5", "conv", (chans[4], out, 5, 1, 0)))
    elif enc == "unet":                                    # default.py:85-104,213-239
        specs.append((p + "conv_layers.0", "conv", (1, f, 160, 4, 79)))
        specs.append((p + "conv_layers.1", "conv", (f, 2 * f, 4, 2, 1)))
        specs.append((p + "conv_layers.2", "conv", (2 * f, 4 * f, 4, 2, 1)))
        for i, c in enumerate((f, 2 * f, 4 * f)):
            specs.append((p + f"activations.{i}.0", "bn", (c,)))
        c = 4 * f
        for i in range(1, 8):
            cin = c if i <= 4 else 2 * c
            specs.append((p + f"ublock.convblock{i}.conv", "conv", (cin, c, 3, 1, 1)))
            specs.append((p + f"ublock.convblock{i}.bn", "bn", (c,)))
        specs.append((p + "fc", "conv", (c, out, 200, 1, 0)))
    else:
        raise ValueError(enc)
    hid = cfg["latent_vector_size"] - cfg["noise_size"]    # default.py:19
    specs.append(("audio_rnn.rnn", "gru", (out, hid, cfg["n_cells"])))
    specs.append(("noise_gen.rnn", "gru", (cfg["noise_size"], cfg["noise_size"], 1)))
    size = cfg["size"]                                     # default.py:146-161
    specs.append(("decoder.fc1", "lin", (cfg["latent_vector_size"], size)))
    specs.append(("decoder.bn1", "bn", (size,)))
    for b in range(cfg["nblocks_gen"]):                    # default.py:171-181
        q = f"decoder.blocks.{b}."
        specs += [(q + "fc1", "lin", (size, size)), (q + "fc2", "lin", (size, size)),
                  (q + "bn1", "bn", (size,)), (q + "bn2", "bn", (size,))]
    specs.append(("decoder.lastfc", "lin", (size, cfg["output_size"])))
    return specs


def _critic_layer_specs(cfg):
    ch, code, cin = cfg["channels"], cfg["code_size"], cfg["output_size"]
    # Q9: the ablated critic does not forward init_ker -> StickDiscriminator default 9
    k0 = 9 if cfg["ablated"] else cfg["init_kernel"]
    specs = [("stick_d.conv1", "conv", (cin, ch, k0, 1, (k0 - 1) // 2))]     # default.py:326-328
    for b in range(2):                                                       # n_blocks=2 default
        specs += [(f"stick_d.blocks.{b}.conv1", "conv", (ch, ch, 7, 1, 3)),
                  (f"stick_d.blocks.{b}.conv2", "conv", (ch, ch, 7, 1, 3))]
    specs.append(("stick_d.fconv", "conv", (ch, code, cfg["stick_length"], 1, 0)))
    if not cfg["ablated"]:                                                   # default.py:294-303
        c = [1, 32, 64, 128, 256, 512]
        for i in range(1, 6):
            specs.append((f"audio_d.l{i}", "conv", (c[i - 1], c[i], 25, 4, 11)))
        specs.append(("audio_d.l6", "conv", (512, code, 75, 1, 0)))
        specs.append(("fc1", "lin", (2 * code, 128)))
    else:
        specs.append(("fc1", "lin", (code, 128)))
    specs.append(("fc2", "lin", (128, 1)))
    return specs


def _materialise(specs):
    """Create torch layers in constructor order (default init draws from the
    global CPU generator), then apply xavier_normal_ in module-traversal order.
    BatchNorm layers draw nothing.  Module traversal order == the order of
    ``specs`` restricted to weight-bearing layers, except that for the default /
    unet encoders ModuleList ``conv_layers`` precedes ``activations`` (already
    reflected in the spec order above)."""
    mods = []
    for name, kind, a in specs:
        if kind == "conv":
            m = nn.Conv1d(a[0], a[1], a[2], stride=a[3], padding=a[4])
        elif kind == "lin":
            m = nn.Linear(a[0], a[1])
        elif kind == "gru":
            m = nn.GRU(a[0], a[1], a[2], batch_first=True)
        elif kind == "bn":
            m = nn.BatchNorm1d(a[0])
        mods.append((name, kind, m))
    for name, kind, m in mods:
        if kind in ("conv", "lin"):
            nn.init.xavier_normal_(m.weight)
        elif kind == "gru":
            for pname, p in m.named_parameters():
                if "weight" in pname:
                    nn.init.xavier_normal_(p)
    P = OrderedDict()
    for name, kind, m in mods:
        for k, v in m.state_dict().items():
            P[f"{name}.{k}"] = v.detach().clone()
    return P


def creation_order_specs_generator(cfg):
    """Creation order differs from traversal order only for BN placement, which
    draws no random numbers, so a single ordered list serves both."""
    return _gen_layer_specs(cfg)


def init_generator_params(cfg):
    return _materialise(_gen_layer_specs(cfg))


def init_critic_params(cfg):
    return _materialise(_critic_layer_specs(cfg))


def pre_bn_bias_names(P):
    """Biases of layers that feed a train-mode BatchNorm.  Their true gradient is
    exactly zero (BN removes any per-channel shift), so every implementation —
    the reference included — produces rounding noise there, and Adam then turns
    that noise into +-lr random walks.  They cannot influence any output; parity
    checks treat their gradients with an absolute tolerance and skip their
    post-Adam values."""
    out = []
    keys = list(P.keys())
    for k in keys:
        if not k.endswith(".bias"):
            continue
        base = k[:-5]
        m = re.match(r"audio_enc\.model\.conv_layers\.(\d+)$", base)
        if m and f"audio_enc.model.activations.{m.group(1)}.0.weight" in P:
            out.append(k)
        m = re.match(r"audio_enc\.model\.l(\d)$", base)
        if m and f"audio_enc.model.bn{m.group(1)}.weight" in P:
            out.append(k)
        if re.match(r"audio_enc\.model\.ublock\.convblock\d\.conv$", base):
            out.append(k)
        if base == "decoder.fc1" or re.match(r"decoder\.blocks\.\d+\.fc2$", base):
            out.append(k)
    return out


def trainable_names(P):
    return [k for k in P if not (k.endswith("running_mean") or k.endswith("running_var")
                                 or k.endswith("num_batches_tracked"))]


# --------------------------------------------------------------------------
# audio windowing (utils.py:329-353): zero-pad pad//2 left, pad-pad//2 right,
# then windows of `win` samples every `stride` samples.  Pure indexing.
# --------------------------------------------------------------------------

def slice_audio_batch(audio, win, stride, pad):
    one = audio.dim() == 1
    a = audio.unsqueeze(0) if one else audio
    a = F.pad(a, (pad // 2, pad - pad // 2))
    n = (a.shape[-1] - win) // stride + 1
    out = a.unfold(-1, win, stride)[:, :n].contiguous()
    return out[0] if one else out


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------

def _conv(P, name, x, stride=1, pad=0):
    return F.conv1d(x, P[name + ".weight"], P[name + ".bias"], stride=stride, padding=pad)


def _bn(P, name, x, train):
    """nn.BatchNorm1d, eps 1e-5, momentum 0.1; train mode normalises with the
    biased batch variance and updates running stats with the unbiased one."""
    if train:
        P[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, P[name + ".running_mean"], P[name + ".running_var"],
                        P[name + ".weight"], P[name + ".bias"], training=train,
                        momentum=0.1, eps=1e-5)


_TAPS = None        # set by activation_margins(): list of (rows per channel, min|x| / rms(x))


def _tap(x):
    if _TAPS is not None:
        with torch.no_grad():
            rms = float(x.pow(2).mean().sqrt())
            _TAPS.append((x.numel() // x.shape[1], float(x.abs().min()) / max(rms, 1e-30)))


def _relu(x):
    _tap(x)
    return F.relu(x)


def _lrelu(x):
    _tap(x)
    return F.leaky_relu(x, 0.2)


class activation_margins:
    """Context manager for the parity tests: records, for every ReLU / LeakyReLU input
    evaluated inside it, (rows per channel, smallest |pre-activation| relative to the layer's
    rms).  Gradients are discontinuous at the kink, so test points are chosen where no SMALL
    layer (few rows per channel: one flipped unit moves a gradient by ~1/rows) has a unit
    within rounding noise of zero."""

    def __enter__(self):
        global _TAPS
        _TAPS = []
        return self

    def __exit__(self, *a):
        global _TAPS
        self.taps, _TAPS = _TAPS, None

    def margin(self, max_rows):
        m = [r for n, r in self.taps if n <= max_rows]
        return min(m) if m else float("inf")


def _final_activ(x, activ):
    if activ == "relu":
        return _relu(x)
    if activ == "tanh":
        return torch.tanh(x)
    return x


def encoder_default(P, x, activ, train):
    """default.py:59-82.  x (N,1,3200) -> (N,250)."""
    p = "audio_enc.model."
    x = _relu(_bn(P, p + "activations.0.0", _conv(P, p + "conv_layers.0", x, 50, 124), train))
    for i in range(1, 6):
        x = _relu(_bn(P, p + f"activations.{i}.0", _conv(P, p + f"conv_layers.{i}", x, 2, 1), train))
    x = _final_activ(_conv(P, p + "conv_layers.6", x), activ)
    return x.squeeze()                                      # Q10


def encoder_wavegan(P, x, activ, train):
    """default.py:114-143.  lengths 3200->794->193->43->5->1."""
    p = "audio_enc.model."
    for i in range(1, 5):
        x = _relu(_bn(P, p + f"bn{i}", _conv(P, p + f"l{i}", x, 4, 0), train))
    return _final_activ(_conv(P, p + "l5", x), activ).squeeze(-1)


def encoder_unet(P, x, activ, train):
    """default.py:85-111,213-246."""
    p = "audio_enc.model."
    lr = lambda t: _lrelu(t)
    x = lr(_bn(P, p + "activations.0.0", _conv(P, p + "conv_layers.0", x, 4, 79), train))
    x = lr(_bn(P, p + "activations.1.0", _conv(P, p + "conv_layers.1", x, 2, 1), train))
    x = lr(_bn(P, p + "activations.2.0", _conv(P, p + "conv_layers.2", x, 2, 1), train))

    def blk(i, t):
        q = p + f"ublock.convblock{i}"
        return lr(_bn(P, q + ".bn", _conv(P, q + ".conv", t, 1, 1), train))

    down = lambda t: F.max_pool1d(t, 2, 2)
    up = lambda t: F.interpolate(t, scale_factor=2, mode="linear", align_corners=False)
    x1 = blk(1, x)
    x2 = blk(2, down(x1))
    x3 = blk(3, down(x2))
    x4 = blk(4, down(x3))
    x3 = blk(5, torch.cat((up(x4), x3), 1))
    x2 = blk(6, torch.cat((up(x3), x2), 1))
    x = blk(7, torch.cat((up(x2), x1), 1))
    return _final_activ(_conv(P, p + "fc", x), activ).squeeze()


ENCODERS = {"default": encoder_default, "wavegan": encoder_wavegan, "unet": encoder_unet}


def gru_forward(P, name, x, n_layers):
    """torch.nn.GRU (batch_first, h0 = 0), gate rows ordered [r|z|n]:
        r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)
        z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
        n = tanh(W_in x + b_in + r * (W_hn h + b_hn))
        h' = (1 - z) * n + z * h
    x (B,T,I) -> (B,T,H).  Equal-length PackedSequence == dense (Q15)."""
    B, T, _ = x.shape
    for l in range(n_layers):
        w_ih, w_hh = P[f"{name}.weight_ih_l{l}"], P[f"{name}.weight_hh_l{l}"]
        b_ih, b_hh = P[f"{name}.bias_ih_l{l}"], P[f"{name}.bias_hh_l{l}"]
        H = w_hh.shape[1]
        gi = x @ w_ih.t() + b_ih
        h = x.new_zeros(B, H)
        outs = []
        for t in range(T):
            gh = h @ w_hh.t() + b_hh
            r = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
            outs.append(h)
        x = torch.stack(outs, 1)
    return x


def decoder_forward(P, x, nblocks, train):
    """default.py:146-192.  Q1: LinearBlock computes x + relu(bn2(fc2(x))); the
    fc1->bn1 branch is dead but bn1 still updates its running statistics."""
    x = _relu(_bn(P, "decoder.bn1", F.linear(x, P["decoder.fc1.weight"], P["decoder.fc1.bias"]), train))
    for b in range(nblocks):
        q = f"decoder.blocks.{b}."
        _bn(P, q + "bn1", F.linear(x, P[q + "fc1.weight"], P[q + "fc1.bias"]), train)   # dead branch
        x = x + _relu(_bn(P, q + "bn2", F.linear(x, P[q + "fc2.weight"], P[q + "fc2.bias"]), train))
    return F.linear(x, P["decoder.lastfc.weight"], P["decoder.lastfc.bias"])


def generator_forward(P, cfg, slices, noise, train=True):
    """default.py:25-42.  slices (B,T,W), noise (B,T,noise_size) -> (B*T, output_size)."""
    B, T, W = slices.shape
    x = ENCODERS[cfg["enc_type"]](P, slices.reshape(B * T, 1, W), cfg["activ"], train)
    x = x.view(B, T, cfg["input_vector_size"])
    x = gru_forward(P, "audio_rnn.rnn", x, cfg["n_cells"])
    n = gru_forward(P, "noise_gen.rnn", noise, 1)
    x = torch.cat((x, n), -1).reshape(B * T, cfg["latent_vector_size"])
    return decoder_forward(P, x, cfg["nblocks_gen"], train)


def stick_d_forward(P, x, cfg):
    """default.py:322-346,195-210.  x (B,69,T) -> (B,code)."""
    k0 = P["stick_d.conv1.weight"].shape[-1]
    x = _relu(_conv(P, "stick_d.conv1", x, 1, (k0 - 1) // 2))
    for b in range(2):
        y = _relu(_conv(P, f"stick_d.blocks.{b}.conv1", x, 1, 3))
        y = _relu(_conv(P, f"stick_d.blocks.{b}.conv2", y, 1, 3))
        x = x + y
    return _final_activ(_conv(P, "stick_d.fconv", x), cfg["activ"]).squeeze(-1)


def audio_d_forward(P, c, cfg):
    """default.py:294-319.  c (B,1,A) -> (B,code)."""
    for i in range(1, 6):
        c = _relu(_conv(P, f"audio_d.l{i}", c, 4, 11))
    return _final_activ(_conv(P, "audio_d.l6", c), cfg["activ"]).squeeze(-1)


def critic_forward(P, cfg, x, c=None):
    """default.py:263-270 (full) / :286-291 (ablated).  -> (B,1)."""
    s = stick_d_forward(P, x, cfg)
    if not cfg["ablated"]:
        s = torch.cat((s, audio_d_forward(P, c, cfg)), -1)
    h = _relu(F.linear(s, P["fc1.weight"], P["fc1.bias"]))
    return F.linear(h, P["fc2.weight"], P["fc2.bias"])


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------

def gradient_penalty(P, cfg, real, fake, audio, alpha):
    """losses.py:5-60 with is_seq=True, lp=False.  `alpha` (B,1) is the
    U[0,1) draw the reference takes from the CPU generator (Q3).
    Returns (gp, g0, g1): the scalar (differentiable w.r.t. P) and the
    per-sample input gradients (detached; g1 None when ablated)."""
    B = real.shape[0]
    a = alpha.view(B, 1)
    inter = (a * real.detach().reshape(B, -1) + (1 - a) * fake.detach().reshape(B, -1))
    inter = inter.view(B, cfg["output_size"], -1).requires_grad_(True)
    if cfg["ablated"]:
        out = critic_forward(P, cfg, inter)
        (g0,) = torch.autograd.grad(out, inter, torch.ones_like(out), create_graph=True)
        n0 = torch.sqrt((g0.reshape(B, -1) ** 2).sum(1) + 1e-12)
        return ((n0 - 1) ** 2).mean(), g0.detach(), None
    aud = audio.detach().requires_grad_(True)
    out = critic_forward(P, cfg, inter, aud)
    g0, g1 = torch.autograd.grad(out, (inter, aud), torch.ones_like(out), create_graph=True)
    n0 = torch.sqrt((g0.reshape(B, -1) ** 2).sum(1) + 1e-12)
    n1 = torch.sqrt((g1.reshape(B, -1) ** 2).sum(1) + 1e-12)
    return ((n0 - 1) ** 2).mean() + ((n1 - 1) ** 2).mean(), g0.detach(), g1.detach()


def jerkiness(seq):
    """losses.py:85-89 for (B, C, T) input."""
    d = seq[:, :, 3:] - 3 * seq[:, :, 2:-1] + 3 * seq[:, :, 1:-2] - seq[:, :, :-3]
    return (d ** 2).sum(dim=1).mean()


def validation_l1(G, cfg, real_bt, audio, noise):
    """phase3/train.py:245-261 for one validation batch: eval-mode generator, mean |real - fake|.
    Returns (l1, fake (B, 69, T))."""
    B, T, Oo = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    with torch.no_grad():
        sl = slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"], cfg["pad_samples"])
        real = real_bt.reshape(B, T, Oo).permute(0, 2, 1)
        fake = generator_forward(G, cfg, sl, noise, train=False).view(B, T, Oo).permute(0, 2, 1)
        return float((real - fake).abs().mean()), fake


def tv_loss(seq):
    """losses.py:76-82: mean |x[t+1]-x[t]| over (B,C,T-1)."""
    return (seq[:, :, 1:] - seq[:, :, :-1]).abs().mean()


# --------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults, train.py:102-103): betas (0.9, 0.999),
# eps 1e-8, no weight decay; parameters without a gradient are skipped (Q1).
# --------------------------------------------------------------------------

class AdamState:
    def __init__(self, P, lr):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, 0.9, 0.999, 1e-8, {}
        self.m = {k: torch.zeros_like(P[k]) for k in trainable_names(P)}
        self.v = {k: torch.zeros_like(P[k]) for k in trainable_names(P)}

    def step(self, P, grads):
        for k, g in grads.items():
            if g is None:
                continue
            t = self.t[k] = self.t.get(k, 0) + 1
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1, bc2 = 1 - self.b1 ** t, 1 - self.b2 ** t
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            P[k].addcdiv_(self.m[k], denom, value=-(self.lr / bc1))


# --------------------------------------------------------------------------
# train-step body (phase3/train.py:186-237)
# --------------------------------------------------------------------------

def _leaf(P):
    """Differentiable view of the trainable entries of P."""
    Q = OrderedDict()
    for k, v in P.items():
        Q[k] = v.requires_grad_(True) if v.is_floating_point() and k in trainable_names(P) else v
    return Q


def critic_iteration(G, D, cfg, real_bt, audio, noise, alpha, adam_d=None):
    """One critic update (train.py:187-216).

    real_bt (B,T,23,3) or (B,T,69); audio (B,A); noise (B,T,noise_size);
    alpha (B,1).  Mutates G's BN running stats (Q2) and, if `adam_d` is given,
    D's parameters.  Returns scalars + critic grads + the generated poses."""
    B, T, O = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    slices = slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"], cfg["pad_samples"])
    with torch.no_grad():
        fake = generator_forward(G, cfg, slices, noise, train=True)
    fake = fake.view(B, T, O).permute(0, 2, 1).contiguous()
    real = real_bt.reshape(B, T, O).permute(0, 2, 1).contiguous()
    aud = audio.unsqueeze(1)
    for v in D.values():
        v.grad = None
    Dl = _leaf(D)
    gp, g0, g1 = gradient_penalty(Dl, cfg, real, fake, None if cfg["ablated"] else aud, alpha)
    if cfg["ablated"]:
        err_real = critic_forward(Dl, cfg, real).mean()
        err_fake = critic_forward(Dl, cfg, fake).mean()
    else:
        err_real = critic_forward(Dl, cfg, real, aud).mean()
        err_fake = critic_forward(Dl, cfg, fake, aud).mean()
    err = err_fake - err_real + cfg["gamma"] * gp
    names = trainable_names(D)
    gl = torch.autograd.grad(err, [Dl[k] for k in names], allow_unused=True)
    grads = OrderedDict((k, g) for k, g in zip(names, gl))
    for v in D.values():
        v.requires_grad_(False) if v.is_floating_point() else None
    out = dict(loss_critic=float(err), gp=float(gp), w_dist=float(err_fake - err_real),
               err_real=float(err_real), err_fake=float(err_fake),
               fake=fake.detach(), g0=g0, g1=g1, grads=grads)
    if adam_d is not None:
        with torch.no_grad():
            adam_d.step(D, grads)
    return out


def generator_update(G, D, cfg, real_bt, audio, noise, adam_g=None):
    """Generator update (train.py:222-237): L = E[D(real)] - E[D(fake)] + beta*L1 + eta*TV."""
    B, T, O = real_bt.shape[0], cfg["stick_length"], cfg["output_size"]
    slices = slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"], cfg["pad_samples"])
    real = real_bt.reshape(B, T, O).permute(0, 2, 1).contiguous()
    aud = audio.unsqueeze(1)
    Gl = _leaf(G)
    fake = generator_forward(Gl, cfg, slices, noise, train=True).view(B, T, O).permute(0, 2, 1)
    l1 = (real - fake).abs().mean()
    with torch.no_grad():
        err_real = (critic_forward(D, cfg, real) if cfg["ablated"] else critic_forward(D, cfg, real, aud)).mean()
    err_fake = (critic_forward(D, cfg, fake) if cfg["ablated"] else critic_forward(D, cfg, fake, aud)).mean()
    tv = tv_loss(fake)
    err = err_real - err_fake + cfg["beta"] * l1 + cfg["eta"] * tv
    names = trainable_names(G)
    gl = torch.autograd.grad(err, [Gl[k] for k in names], allow_unused=True)
    grads = OrderedDict((k, g) for k, g in zip(names, gl))
    for v in G.values():
        v.requires_grad_(False) if v.is_floating_point() else None
    out = dict(loss_gen=float(err), l1=float(l1), tv=float(tv), err_real=float(err_real),
               err_fake=float(err_fake), fake=fake.detach(), grads=grads)
    if adam_g is not None:
        with torch.no_grad():
            adam_g.step(G, grads)
    return out


# --------------------------------------------------------------------------
# synthetic data of the reference shape
# --------------------------------------------------------------------------

def synthetic_batch(cfg, B, seed):
    """real (B,T,23,3) ~ U[0,1); audio (B,A) ~ 0.3*U(-1,1); noise ~ N(0,1);
    alpha ~ U[0,1).  Drawn from a private generator in this fixed order."""
    g = torch.Generator().manual_seed(seed)
    T, A = cfg["stick_length"], cfg["audio_length"]
    real = torch.rand(B, T, 23, 3, generator=g)
    audio = (torch.rand(B, A, generator=g) * 2 - 1) * 0.3
    noise = torch.randn(B, T, cfg["noise_size"], generator=g)
    alpha = torch.rand(B, 1, generator=g)
    noise_g = torch.randn(B, T, cfg["noise_size"], generator=g)
    return real, audio, noise, alpha, noise_g


def train_step(G, D, cfg, B, step_idx, adam_g, adam_d, seed0=1234):
    """One full train step: n_critic critic iterations (fresh batch each, Q7)
    followed by a generator update on the last batch."""
    logs = []
    n = cfg["n_critic_steps"]
    for i in range(n):
        real, audio, noise, alpha, noise_g = synthetic_batch(cfg, B, seed0 + step_idx * n + i)
        logs.append(critic_iteration(G, D, cfg, real, audio, noise, alpha, adam_d))
    logs.append(generator_update(G, D, cfg, real, audio, noise_g, adam_g))
    return logs


# --------------------------------------------------------------------------
# helpers shared by the fixture generator and the parity tests
# --------------------------------------------------------------------------

def perturb_params(P, scale=0.05):
    """Deterministic, RNG-free perturbation giving a non-initial state: BN
    gamma/beta away from (1,0), biases away from their init, running stats moved.
    Reproducible bit-for-bit anywhere (pure float64 arithmetic on arange)."""
    for i, (k, v) in enumerate(P.items()):
        if not v.is_floating_point():
            continue
        idx = torch.arange(v.numel(), dtype=torch.float64)
        w = torch.cos(idx * 0.7310585786 + 1.6180339887 * (i + 1)).view_as(v)
        amp = scale * float(v.abs().mean()) if v.numel() > 1 else scale
        amp = max(amp, scale * 0.1)
        if k.endswith("running_var"):
            v.copy_((v.double() * (1.0 + 0.3 * w)).float())
        else:
            v.add_((amp * w).float())
    return P


def tensor_digest(t, nsamp=64):
    """(sum, l2, maxabs) in float64 plus `nsamp` evenly strided elements."""
    f = t.detach().double().reshape(-1)
    step = max(1, f.numel() // nsamp)
    return dict(sum=float(f.sum()), l2=float(f.norm()), maxabs=float(f.abs().max()),
                samples=f[::step][:nsamp].float().clone(), n=int(f.numel()))
