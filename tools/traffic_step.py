#!/usr/bin/env python
"""One eager, single-stream train step bracketed by cudaProfilerStart / Stop with an NVTX range per kernel family, for
    ncu --profile-from-start off --nvtx --nvtx-include "<family>/" ...      (tools/gpu_profiles.sh)
so that the DRAM traffic ncu reports belongs to exactly the launches bench.py's roofline counts for that family.
Writes the per-family call counts of the profiled step to gpurun_out/traffic_counts_b<batch>.json.
    python tools/traffic_step.py [batch] [enc] [gemm]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import config as O, ops                                               # noqa: E402
from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator       # noqa: E402
from music2dance_b200.trainer import Phase3Trainer                                          # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 7
enc = sys.argv[2] if len(sys.argv) > 2 else "default"
gemm = sys.argv[3] if len(sys.argv) > 3 else "tf32x3"
dev = "cuda:0"
cfg = O.make_cfg(enc_type=enc)
nc = cfg["n_critic_steps"]
ops.set_gemm_mode(gemm)
torch.manual_seed(0)
gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"],
                        cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"], cfg["enc_type"],
                        cfg["activ"], dev)
critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                               init_ker=cfg["init_kernel"], activ=cfg["activ"], device=dev)
tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=False)
tr.overlap, tr.D.par, tr.G.par = False, False, False          # one stream, like bench.py's instrumented step
bs = [O.synthetic_batch(cfg, B, 1234 + i) for i in range(nc)]
tr.load_batches(*[torch.stack([b[j] for b in bs]) for j in range(4)], bs[-1][4])
for _ in range(2):
    tr.train_step()
torch.cuda.synchronize()

counts = {}
FAM = {"rowconv": lambda a, k: "rowconv_c1" if k["Cc"] == 1 else "rowconv",
       "wgrad": lambda a, k: "wgrad_c1" if k["Cc"] == 1 else "wgrad",
       "gru_forward": lambda a, k: "gru_forward", "gru_backward": lambda a, k: "gru_backward",
       "adam_pack": lambda a, k: "adam", "conv_dgrad_c1": lambda a, k: "conv_dgrad_c1"}
for name, fam_fn in FAM.items():
    orig = getattr(ops, name)

    def wrapped(*a, _o=orig, _f=fam_fn, **k):
        fam = _f(a, k)
        counts[fam] = counts.get(fam, 0) + 1
        torch.cuda.nvtx.range_push(fam)
        try:
            return _o(*a, **k)
        finally:
            torch.cuda.nvtx.range_pop()
    setattr(ops, name, wrapped)

torch.cuda.profiler.start()
tr.train_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/traffic_counts_b{B}.json", "w") as f:
    json.dump({"batch": B, "enc": enc, "gemm": gemm, "calls_per_step": counts}, f)
print(counts)
