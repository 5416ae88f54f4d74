#!/usr/bin/env python
"""Per-kernel-family roofline table from a bench.py line (its `kernel_families` object: CUDA-event time,
algorithmic FLOPs and bytes of every launch of one instrumented eager train step) against the measured peaks the
same line used.  A family is tensor-bound when its algorithmic intensity exceeds 100 FLOP/B (bench.py's rule),
HBM-bound otherwise; the GRU recurrence is latency-bound (120 dependent steps) and is listed without a fraction.
    python tools/roofline_table.py profiles/r01t_bench_b7.json [MEASURED_PEAKS.json]
"""
import json
import sys


def main(path, peaks_path=None):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    pk = {"hbm_gbs": 6420.7, "bf16_tflops_sustained": 1367.5}
    if peaks_path:
        pk.update(json.load(open(peaks_path)))
    tf32_peak = d.get("roofline", {}).get("peak", pk["bf16_tflops_sustained"] / 2.0)
    fam = d["kernel_families"]
    print(f"| family | launches | ms | share of eager step | achieved | bound | peak | fraction |")
    print("|---|---:|---:|---:|---:|---|---:|---:|")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        inten = v["flops"] / max(v["bytes"], 1.0)
        if k.startswith("gru"):
            bound, ach, peak, frac = "latency (recurrence)", f"{v['tflops']:.2f} TFLOP/s", "—", "—"
        elif inten > 100.0:
            bound, ach, peak = "tensor (TF32 dense)", f"{v['tflops']:.1f} TFLOP/s", f"{tf32_peak:.0f}"
            frac = f"{v['tflops'] / tf32_peak:.3f}"
        else:
            bound, ach, peak = "HBM", f"{v['gbs']:.0f} GB/s", f"{pk['hbm_gbs']:.0f}"
            frac = f"{v['gbs'] / pk['hbm_gbs']:.3f}"
        print(f"| `{k}` | {v['launches']} | {v['ms']:.3f} | {v['share_of_eager_step']:.3f} | {ach} | {bound} | {peak} | {frac} |")


if __name__ == "__main__":
    main(*sys.argv[1:3])
