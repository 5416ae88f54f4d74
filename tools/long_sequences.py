#!/usr/bin/env python
"""Generator inference on long sequences (BASELINE.json configs[4]; phase3/test.py:49,69-75): eval-mode
SequenceGenerator through the drop-in API — slice_audio_batch + gen(slices, [T], noise) — at T = 120 / 750 / 3000
frames (4.8 s / 30 s / 2 min of audio at 25 fps) for the three audio encoders.  CUDA events, 3 warm-up + 10 timed
calls, median; one JSON line per (encoder, T, batch).
    python tools/long_sequences.py [out.jsonl]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import config as C                                                    # noqa: E402
from music2dance_b200.archis.default import SequenceGenerator                               # noqa: E402
from music2dance_b200.utils import slice_audio_batch                                         # noqa: E402

dev = "cuda:0"
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
for enc in ("default", "wavegan", "unet"):
    cfg = C.make_cfg(enc_type=enc)
    torch.manual_seed(0)
    gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"],
                            cfg["size"], cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"],
                            cfg["enc_type"], cfg["activ"], dev)
    gen.eval()
    for B, T in ((1, 120), (1, 750), (1, 3000), (8, 750)):
        g = torch.Generator().manual_seed(T)
        audio = ((torch.rand(B, T * cfg["cutting_stride"], generator=g) - 0.5) * 0.6).to(dev)
        noise = torch.randn(B, T, cfg["noise_size"], generator=g).to(dev)

        def call():
            with torch.no_grad():
                sl = slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"], cfg["pad_samples"])
                return gen(sl, [T] * B, noise=noise)

        for _ in range(3):
            y = call()
        assert y.shape == (B * T, cfg["output_size"]) and bool(torch.isfinite(y).all())
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        line = dict(workload="phase3 generator inference (eval mode), drop-in API incl. audio windowing",
                    enc=enc, batch=B, frames=T, audio_seconds=T / 25.0, ms_per_call=round(ms, 3),
                    frames_per_s=round(B * T / ms * 1e3, 1), times_realtime=round(B * T / 25.0 / (ms * 1e-3), 1),
                    gemm="tf32x3", timing="CUDA events, median of 10 after 3 warm-up calls, inputs resident in HBM")
        print(json.dumps(line), flush=True)
        if out:
            out.write(json.dumps(line) + "\n")
if out:
    out.close()
