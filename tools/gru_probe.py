#!/usr/bin/env python
"""Device time of the GRU recurrence kernels alone (audio GRU layer: B sequences x 120 steps x H = 240; noise GRU H = 10)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import ops                                                      # noqa: E402

dev = "cuda:0"
T = 120
for B, H in ((7, 240), (7, 10), (64, 240)):
    gi = torch.randn(B * T, 3 * H, device=dev) * 0.3
    w = torch.randn(3 * H, H, device=dev) / H ** 0.5
    b = torch.randn(3 * H, device=dev) * 0.1
    h = torch.empty(B * T, H, device=dev)
    sv = torch.empty(B * T, 4 * H, device=dev)
    ts = []
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gru_forward(gi, w, b, h, H, sv, B, T, H)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"gru_forward B={B} T={T} H={H}: median {ts[len(ts) // 2]:.1f} us ({ts[len(ts) // 2] / T:.2f} us / step), "
          f"spin={os.environ.get('M2D_GRU_SPIN', '0')}", flush=True)
