#!/usr/bin/env python
"""How far do two LEGITIMATE fp32 evaluations of the reference algorithm drift apart over chained Adam steps?
(the un-synced whole-step test compares at TOL_CHAINED = 2e-2; this calibrates that number on the
CPU side alone.)  The oracle's critic iteration is run from identical state on identical batches
  (a) fp32 with all host threads, (b) fp32 with 3 threads (another summation order), (c) fp64 (ground truth),
for `nc` chained iterations with Adam; printed: relative deviation of loss_critic / gp / w_dist per iteration.
    python tools/kink_calibration.py [variant] [B] [nc]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import phase3_oracle as O          # noqa: E402
from tests.parity import VARIANTS              # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "ablated"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nc = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = O.make_cfg(n_critic_steps=nc, **VARIANTS[variant])
torch.manual_seed(0)
G0, D0 = O.init_generator_params(cfg), O.init_critic_params(cfg)
batches = [O.synthetic_batch(cfg, B, 4000 + i) for i in range(nc)]


def run(dtype, threads):
    torch.set_num_threads(threads)
    cast = lambda P: {k: (v.to(dtype).clone() if v.is_floating_point() else v.clone()) for k, v in P.items()}
    G, D = cast(G0), cast(D0)
    ad = O.AdamState(D, cfg["lr_critic"])
    out = []
    for b in batches:
        o = O.critic_iteration(G, D, cfg, b[0].to(dtype), b[1].to(dtype), b[2].to(dtype), b[3].to(dtype), ad)
        out.append({k: o[k] for k in ("loss_critic", "gp", "w_dist")})
    return out


n = os.cpu_count() or 8
ref = run(torch.float64, n)
a = run(torch.float32, n)
b = run(torch.float32, 3)
rel = lambda x, y: abs(x - y) / max(abs(y), 1e-2)
print(f"variant {variant}, batch {B}, {nc} chained critic iterations (Adam lr {cfg['lr_critic']}); relative deviations")
print("| iteration | fp32 (all threads) vs fp64 | fp32 (3 threads) vs fp64 | fp32 vs fp32 (thread count) |")
print("|---|---|---|---|")
for i in range(nc):
    f = lambda u, v: " / ".join(f"{rel(u[i][k], v[i][k]):.1e}" for k in ("loss_critic", "gp", "w_dist"))
    print(f"| {i} | {f(a, ref)} | {f(b, ref)} | {f(a, b)} |")
print("(loss_critic / gp / w_dist; gp values:", ", ".join(f"{r['gp']:.4f}" for r in ref), ")")
