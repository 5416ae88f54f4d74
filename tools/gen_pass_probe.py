#!/usr/bin/env python
"""The batched generator pass of a train step alone (trainer.gen_groups, default: all n_critic iterations in one pass):
    python tools/gen_pass_probe.py [batch]            # CUDA-event time of the pass, eager and as a graph
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gen_pass.csv \
        python tools/gen_pass_probe.py 7 once          # per-kernel launch list of ONE pass
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import config as O                                                    # noqa: E402
from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator       # noqa: E402
from music2dance_b200.trainer import Phase3Trainer                                          # noqa: E402

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 7
once = len(sys.argv) > 2 and sys.argv[2] == "once"
cfg = O.make_cfg(enc_type=os.environ.get("M2D_ENC", "default"))
nc = cfg["n_critic_steps"]
torch.manual_seed(0)
gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"],
                        cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"], cfg["enc_type"],
                        cfg["activ"], dev)
critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                               init_ker=cfg["init_kernel"], activ=cfg["activ"], device=dev)
tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=False)
bs = [O.synthetic_batch(cfg, B, 1234 + i) for i in range(nc)]
tr.load_batches(*[torch.stack([b[j] for b in bs]) for j in range(4)], bs[-1][4])
print("generator pass plan:", tr.gen_groups)
with torch.cuda.device(dev):
    tr._gen_forward(0)                      # allocates the workspace
    torch.cuda.synchronize()
    if once:
        torch.cuda.profiler.start()
        tr._gen_forward(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        sys.exit(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        tr._gen_forward(0)
    e0.record()
    for _ in range(10):
        tr._gen_forward(0)
    e1.record()
    torch.cuda.synchronize()
    print(f"eager: {e0.elapsed_time(e1) / 10 * 1000:.1f} us per pass")
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        tr._gen_forward(0)
    for _ in range(3):
        g.replay()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph: {e0.elapsed_time(e1) / 10 * 1000:.1f} us per pass")
