#!/usr/bin/env python
"""Bring-up probe for the tcgen05 kernels: one GEMM family x one mode per process (a device trap
must not take the other cases down).   python tools/tc_probe.py {fwd|wgrad|timing} {fp32|tf32|tf32x3}"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import ops            # noqa: E402
from music2dance_b200.ops import Mat        # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def fwd(mode):
    ops.set_gemm_mode(mode)
    g = torch.Generator().manual_seed(0)
    ws = torch.empty(1 << 24, device=DEV)
    for (M, N, K) in [(128, 128, 32), (128, 64, 64), (256, 128, 256), (840, 256, 250), (2400, 64, 800),
                      (525, 512, 6400), (7, 100, 38400), (1000, 250, 2048), (300, 72, 96)]:
        x = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / K ** 0.5
        ref = x.double() @ w.double().t()
        X = Mat.of(x.to(DEV), 1, M, K)
        Y = Mat.of(torch.empty(M, N, device=DEV), 1, M, N)
        ops.rowconv(X, w.to(DEV), Y, T=1, Cc=K, N=N, ws=ws)
        torch.cuda.synchronize()
        y = Y.t.view(M, N).cpu().double()
        print(f"fwd {mode} M={M} N={N} K={K}: rel err {rel(y, ref):.3e}  fp32-torch {rel((x @ w.t()).double(), ref):.3e}",
              flush=True)
        if rel(y, ref) > 1e-2:
            print("   y[0,:6]  ", y[0, :6].tolist())
            print("   ref[0,:6]", ref[0, :6].tolist())
            print("   y[1,:6]  ", y[1, :6].tolist())
            print("   ref[1,:6]", ref[1, :6].tolist())


def wgrad(mode):
    ops.set_gemm_mode(mode)
    g = torch.Generator().manual_seed(1)
    ws = torch.empty(1 << 24, device=DEV)
    for (Kt, Cout, Cc) in [(32, 128, 128), (64, 128, 32), (256, 64, 128), (840, 256, 256), (2400, 64, 32),
                           (525, 512, 256), (1000, 100, 200), (4800, 128, 64)]:
        dy = torch.randn(Kt, Cout, generator=g)
        x = torch.randn(Kt, Cc, generator=g)
        ref = dy.double().t() @ x.double()
        dw = torch.zeros(Cout, Cc, device=DEV)
        ops.wgrad(Mat.of(dy.to(DEV), 1, Kt, Cout), Mat.of(x.to(DEV), 1, Kt, Cc), dw, Cout=Cout, T=1, Cc=Cc, ws=ws)
        torch.cuda.synchronize()
        d = dw.cpu().double()
        print(f"wgrad {mode} K={Kt} Cout={Cout} Cc={Cc}: rel err {rel(d, ref):.3e}", flush=True)
        if rel(d, ref) > 1e-2:
            print("   dw[0,:6] ", d[0, :6].tolist())
            print("   ref[0,:6]", ref[0, :6].tolist())
            print("   dw[:6,0] ", d[:6, 0].tolist())
            print("   ref[:6,0]", ref[:6, 0].tolist())


def timing(mode):
    ops.set_gemm_mode(mode)
    ws = torch.empty(1 << 25, device=DEV)
    for (M, N, K) in [(8400, 256, 256), (2100, 512, 6400), (33600, 64, 800), (8400, 128, 1600), (65536, 128, 1024)]:
        X = Mat.of(torch.randn(M, K, device=DEV), 1, M, K)
        w = torch.randn(N, K, device=DEV)
        Y = Mat.of(torch.empty(M, N, device=DEV), 1, M, N)
        for _ in range(3):
            ops.rowconv(X, w, Y, T=1, Cc=K, N=N, ws=ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.rowconv(X, w, Y, T=1, Cc=K, N=N, ws=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"timing fwd {mode} M={M} N={N} K={K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        dw = torch.zeros(N, K, device=DEV)
        DY = Mat.of(torch.randn(M, N, device=DEV), 1, M, N)
        for _ in range(3):
            ops.wgrad(DY, X, dw, Cout=N, T=1, Cc=K, ws=ws)
        e0.record()
        for _ in range(20):
            ops.wgrad(DY, X, dw, Cout=N, T=1, Cc=K, ws=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"timing wgrad {mode} K={M} Cout={N} Cc={K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    ops.check_device(0)
    {"fwd": fwd, "wgrad": wgrad, "timing": timing}[sys.argv[1]](sys.argv[2])
