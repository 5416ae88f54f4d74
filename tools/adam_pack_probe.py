#!/usr/bin/env python
"""Device time of the optimiser step of the critic: fused m2d_adam_pack tables (early / late / whole network) against the
round-1 chain (gradient unpack -> flat Adam -> re-layout), each alone on an idle GPU (CUDA events, L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import config as O, ops                                               # noqa: E402
from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator       # noqa: E402
from music2dance_b200.engine import AdamPack                                                # noqa: E402

dev = "cuda:0"
cfg = O.make_cfg()
torch.manual_seed(0)
critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                               init_ker=cfg["init_kernel"], activ=cfg["activ"], device=dev)
gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"],
                        cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"], cfg["enc_type"],
                        cfg["activ"], dev)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)


def timed(name, fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name:58s} median {ts[len(ts) // 2]:8.1f} us   min {ts[0]:8.1f} us", flush=True)


for name, mod in (("critic", critic), ("generator", gen)):
    eng = mod._engine()
    eng.net.pack()
    n = eng.fp.n_live_padded
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    for t in eng.fp.grad_buffers():
        t.normal_().mul_(1e-2)
    whole = AdamPack(eng.fp, eng.net, m, v)
    print(f"{name}: {n} parameters, {whole.n} items, tile {whole.smem_floats * 4 >> 10} KiB")
    timed(f"{name}: m2d_adam_pack, whole network", lambda: whole.step(2e-4))
    if name == "critic":
        late = [eng.net.a_layers[4], eng.net.a_l6]
        e_tab, l_tab = AdamPack(eng.fp, eng.net, m, v, exclude=late), AdamPack(eng.fp, eng.net, m, v, only=late)
        timed("critic: m2d_adam_pack, early table (all but audio_d.l5/l6)", lambda: e_tab.step(2e-4))
        timed("critic: m2d_adam_pack, late table (audio_d.l5/l6)", lambda: l_tab.step(2e-4))
    timed(f"{name}: round-1 chain unpack -> adam -> pack_batch", lambda: (eng.net.unpack_grads(),
          ops.adam(eng.fp.flat, eng.fp.grad, m, v, n, step, 2e-4), eng.net.pack()))
    timed(f"{name}:   of which flat adam", lambda: ops.adam(eng.fp.flat, eng.fp.grad, m, v, n, step, 2e-4))
    timed(f"{name}:   of which pack_batch", lambda: eng.net.pack())
