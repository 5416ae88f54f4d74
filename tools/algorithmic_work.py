#!/usr/bin/env python
"""Algorithmic work of the phase3 train step recomputed from LAYER SHAPES ("the build must ship a
script that recomputes all of these from layer shapes rather than copying these constants").

The layer shapes are read from the drop-in modules' own Conv1d / Linear / GRU parameter containers (the reference's
module tree, music2dance_b200/archis/default.py) instantiated on the CPU; only the topology — which length each
layer runs at — is written down here, with the reference lines it follows.  Prints, per audio encoder:
  G  generator forward MACs per 120-frame sequence      (phase3/archis/default.py:25-42)
  A  critic audio branch                                  (default.py:294-319)
  S  critic pose branch                                   (default.py:322-346)
  F  fusion MLP                                           (default.py:258-259,268-270)
and the de-duplicated minimum per train step per sequence, 8 (G + 6A + 10S) + 3G + A + 3S = 11G + 49A + 83S MACs
(audio branch once per critic iteration, no all-zero double-backward passes, no unused weight gradients),
plus the HBM-side figures (Adam 28 B/param, gradient all-reduce bytes).
    python tools/algorithmic_work.py            # table
    python tools/algorithmic_work.py --json     # one JSON object
"""
import json
import os
import sys

import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import config as C                                                    # noqa: E402
from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator       # noqa: E402


def sub(m, *names):
    """Child module by (dotted) name — the drop-in containers are plain nn.Module holders, not ModuleLists."""
    for n in names:
        m = m._modules[str(n)]
    return m


def children(m):
    return list(m._modules.values())


def out_len(L, m):
    return (L + 2 * m.padding[0] - m.kernel_size[0]) // m.stride[0] + 1


def conv_macs(m, L):
    """MACs of nn.Conv1d `m` on an input of length L -> (macs, output length)."""
    Lo = out_len(L, m)
    return Lo * m.out_channels * m.in_channels * m.kernel_size[0], Lo


def chain(convs, L):
    total = 0
    for m in convs:
        macs, L = conv_macs(m, L)
        total += macs
    return total, L


def gru_macs(rnn):
    """per time step: 3H(I_l + H) for every layer (input projection + recurrence)."""
    H = rnn.hidden_size
    return sum(3 * H * ((rnn.input_size if l == 0 else H) + H) for l in range(rnn.num_layers))


def encoder_macs(gen, W):
    """per audio window of W samples."""
    model = gen.audio_enc.model
    if gen.enc_type == "default":                              # default.py:59-82: seven convolutions in sequence
        macs, L = chain([sub(model, "conv_layers", i) for i in range(7)], W)
    elif gen.enc_type == "wavegan":                            # default.py:114-143
        macs, L = chain([getattr(model, f"l{i}") for i in range(1, 6)], W)
    else:                                                      # default.py:85-111 + UBlock :213-246
        macs, L = chain([sub(model, "conv_layers", i) for i in range(3)], W)
        ub = model.ublock
        cb = lambda i: getattr(ub, f"convblock{i}").conv
        # encoder path 1-4 at L, L/2, L/4, L/8 (MaxPool1d(2,2) in between); decoder path 5-7 at L/4, L/2, L after
        # Upsample(x2) + skip concatenation (k = 3, padding 1: lengths are preserved)
        for i, l in zip(range(1, 8), (L, L // 2, L // 4, L // 8, L // 4, L // 2, L)):
            m, lo = conv_macs(cb(i), l)
            assert lo == l
            macs += m
        m, L = conv_macs(model.fc, L)
        macs += m
    assert L == 1, (gen.enc_type, L)
    return macs


def generator_macs(gen, T, W):
    dec = gen.decoder
    lin = lambda m: m.in_features * m.out_features
    decoder = lin(dec.fc1) + sum(lin(b.fc1) + lin(b.fc2) for b in children(dec.blocks)) + lin(dec.lastfc)   # incl. the dead fc1 (Q1)
    per_frame = encoder_macs(gen, W) + gru_macs(gen.audio_rnn.rnn) + gru_macs(gen.noise_gen.rnn) + decoder
    return T * per_frame


def critic_macs(critic, T, A):
    sd = critic.stick_d
    pose, L = conv_macs(sd.conv1, T)
    for blk in children(sd.blocks):
        for m in (blk.conv1, blk.conv2):
            mm, L = conv_macs(m, L)
            pose += mm
    mm, L = conv_macs(sd.fconv, L)
    assert L == 1
    pose += mm
    ad = critic.audio_d
    audio, L = chain([getattr(ad, f"l{i}") for i in range(1, 7)], A)
    assert L == 1
    fusion = critic.fc1.in_features * critic.fc1.out_features + critic.fc2.in_features * critic.fc2.out_features
    return audio, pose, fusion


def nparams(m, live_only=False):
    import re
    dead = re.compile(r"decoder\.blocks\.\d+\.(fc1|bn1)\.(weight|bias)$")
    return sum(p.numel() for n, p in m.named_parameters() if not (live_only and dead.search(n)))


def work(enc_type="default"):
    cfg = C.make_cfg(enc_type=enc_type)
    T, W, A, nc = cfg["stick_length"], cfg["audio_feat_samples"], cfg["audio_length"], cfg["n_critic_steps"]
    gen = SequenceGenerator(W, cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"], cfg["output_size"],
                            cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"], enc_type, cfg["activ"], "cpu")
    critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["code_size"], T,
                                   init_ker=cfg["init_kernel"], activ=cfg["activ"], device="cpu")
    G = generator_macs(gen, T, W)
    Aa, S, F = critic_macs(critic, T, A)
    # critic iteration: gen fwd; audio: fwd, dgrad, 2 double-backward passes, wgrad + dgrad; pose: 3 fwd, 1 dgrad,
    # 2 double-backward, 2 x (wgrad + dgrad).  generator update: gen fwd + bwd (dgrad + wgrad), critic audio fwd,
    # pose 2 fwd + 1 dgrad.
    step = nc * (G + 6 * Aa + 10 * S) + 3 * G + Aa + 3 * S
    pD, pG_live = nparams(critic), nparams(gen, live_only=True)
    return dict(enc_type=enc_type, frames=T, n_critic=nc,
                G_macs=G, A_macs=Aa, S_macs=S, F_macs=F,
                step_macs_per_sequence=step, step_gflop_per_sequence=2 * step / 1e9,
                reference_executed_macs_per_sequence=nc * (G + 12 * (Aa + S + F)) + 3 * G + 6 * (Aa + S + F),
                critic_params=pD, generator_params=nparams(gen), generator_live_params=pG_live,
                adam_bytes_per_step=28 * (nc * pD + pG_live),
                allreduce_bytes_per_step=4 * (nc * pD + pG_live))


if __name__ == "__main__":
    rows = [work(e) for e in ("default", "wavegan", "unet")]
    if "--json" in sys.argv:
        print(json.dumps(rows))
    else:
        for r in rows:
            print(f"{r['enc_type']:8s} G {r['G_macs'] / 1e6:9.2f} M  A {r['A_macs'] / 1e6:8.2f} M  S {r['S_macs'] / 1e6:6.2f} M  "
                  f"F {r['F_macs'] / 1e6:5.3f} M | step 11G+49A+83S = {r['step_macs_per_sequence'] / 1e9:7.2f} GMAC = "
                  f"{r['step_gflop_per_sequence']:6.1f} GFLOP / sequence | reference executes "
                  f"{r['reference_executed_macs_per_sequence'] / 1e9:6.1f} GMAC | Adam {r['adam_bytes_per_step'] / 1e9:.2f} GB, "
                  f"all-reduce {r['allreduce_bytes_per_step'] / 1e6:.0f} MB per step")
