#!/usr/bin/env python
"""Operand-split schemes for fp32-grade GEMMs on the tensor pipe, emulated on the CPU.

For a dot product of length K the schemes differ only in how the fp32 operands are represented for the MMAs; the
accumulation is emulated in fp64 so that the table isolates the OPERAND error (the TMEM accumulator adds its own
~1e-6, measured on the GPU, DESIGN.md "Precision").
  tf32            a_hi * b_hi                                   1 TF32 MMA per K-slice (single pass)
  tf32x3          a_hi*b_hi + a_hi*b_lo + a_lo*b_hi             3 TF32 MMAs            (what the kernels issue today)
  tf32+2bf16      a_hi*b_hi (TF32) + bf16(a_hi)*bf16(b_lo) + bf16(a_lo)*bf16(b_hi)
                                                                1 TF32 + 2 BF16 MMAs = 2 TF32-equivalents
hi = round-to-nearest TF32 (10 explicit mantissa bits), lo = a - hi.  Shapes: the critic's audio layers at the
reference batch (K = taps * C_in) with activations ~ ReLU(N(0,1)) and Xavier-scale weights.
    python tools/split_precision_study.py
"""
import torch


def to_tf32(x):
    """round-to-nearest-even to 10 mantissa bits (what cvt.rna.tf32.f32 / the kernels' integer rounding produce)."""
    i = x.contiguous().view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return r.view(torch.float32)


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def study(M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.relu(torch.randn(M, K, generator=g))                       # post-ReLU activations
    b = torch.randn(N, K, generator=g) * (2.0 / (K + N)) ** 0.5          # xavier-scale weights
    ref = a.double() @ b.double().t()
    a_hi, b_hi = to_tf32(a), to_tf32(b)
    a_lo, b_lo = a - a_hi, b - b_hi
    d = lambda x, y: x.double() @ y.double().t()
    out = {
        "tf32": d(a_hi, b_hi),
        "tf32x3": d(a_hi, b_hi) + d(a_hi, to_tf32(b_lo)) + d(to_tf32(a_lo), b_hi),
        "tf32+2bf16": d(a_hi, b_hi) + d(bf16(a_hi), bf16(b_lo)) + d(bf16(a_lo), bf16(b_hi)),
        "fp32 (sequential fp32 sum)": (a @ b.t()).double(),
    }
    scale = float(ref.abs().max())
    return {k: float((v - ref).abs().max()) / scale for k, v in out.items()}


if __name__ == "__main__":
    torch.set_num_threads(8)
    shapes = [("audio_d.l2", 512, 64, 25 * 32), ("audio_d.l3", 512, 128, 25 * 64), ("audio_d.l4", 512, 256, 25 * 128),
              ("audio_d.l5", 256, 512, 25 * 256), ("audio_d.l6", 14, 100, 75 * 512), ("pose conv7", 512, 128, 7 * 128)]
    print(f"{'layer':12s} {'K':>6s} | max |err| / max |y|")
    for name, M, N, K in shapes:
        r = study(M, N, K)
        print(f"{name:12s} {K:6d} | " + "  ".join(f"{k} {v:.2e}" for k, v in r.items()))
