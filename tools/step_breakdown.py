#!/usr/bin/env python
"""Where does a train step spend its time?  Each phase of Phase3Trainer captured in its own CUDA graph and
replayed alone (CUDA events): one generator forward, one critic iteration (without the generator forward,
with / without the optimiser step), the generator update.  The fused step overlaps the nine generator
forwards (side stream) with the eight critic iterations, so its floor is max(9 x gen, 8 x critic) + update.
    python tools/step_breakdown.py [batch]
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
cfg = O.make_cfg()
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
tr.train_step()
torch.cuda.synchronize()
tr.split_pack = False            # every phase below must be a self-contained graph


def timed(name, fn, reps=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:46s} median {ts[len(ts) // 2] * 1e3:9.1f} us   min {ts[0] * 1e3:9.1f} us", flush=True)
    return ts[len(ts) // 2]


def ops_adam():
    """the critic's optimiser step: both m2d_adam_pack tables inline"""
    if tr.apD_late is not None:
        tr.apD_late.step(2e-4)
    tr.apD.step(2e-4)


with torch.cuda.device(dev):
    tg = timed("generator forward (1 of 9)", lambda: tr._gen_forward(0))
    tcn = timed("critic iteration, no optimiser step", lambda: tr.critic_iteration(0, update=False, gen_inline=False))
    tc = timed("critic iteration incl. Adam + re-layout", lambda: tr.critic_iteration(0, update=True, gen_inline=False))
    tu = timed("generator update (critic fwd/bwd, G bwd, Adam)", lambda: tr.generator_update(update=True, gen_inline=False))
    tD = tr.D
    X3 = tD.wk.mat("c:X3", 3 * B, tr.T, tr.O)
    from music2dance_b200.wgan import critic_forward
    tf = timed("critic forward only (3B pose rows, B audio)", lambda: critic_forward(tD, X3, tr.in_audio[0], 3 * B, B, "c", groups=3))
    ta = timed("audio branch forward only", lambda: tD.audio_fwd(tr.in_audio[0], B, "c"))
    tp = timed("weight re-layout alone (pack_batch, init / resume path)", lambda: tD.pack())
    from music2dance_b200.wgan import gradient_penalty_pass, wasserstein_backward
    fw = critic_forward(tD, X3, tr.in_audio[0], 3 * B, B, "c", groups=3)
    timed("Wasserstein backward chain alone", lambda: wasserstein_backward(tD, fw, B, 2 * B, (-1.0, 1.0), B, "c:w", beta=0.0))
    timed("gradient-penalty chain alone (dgrad, tangent, wgrads)",
          lambda: gradient_penalty_pass(tD, fw, B, "c:gp", 10.0, 1.0, tr.gp_buf, tr.k0, tr.k1))
    timed("gradient-penalty dgrad only", lambda: gradient_penalty_pass(tD, fw, B, "c:gp", 10.0, 1.0, tr.gp_buf, tr.k0, tr.k1,
                                                                       weight_grads=False))
    timed("optimiser step (m2d_adam_pack: Adam + re-layouts)", ops_adam)
    print(f"batch {B}: 9 x gen = {9 * tg:.2f} ms, 8 x critic = {8 * tc:.2f} ms, update = {tu:.2f} ms")
