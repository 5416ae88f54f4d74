#!/bin/bash
# TF32 + BF16 cross-term mode (M2D_GEMM=tf32bf16): kernel-level tests, bench lines at batch 7 / 64, parity suite
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_ops_gpu.py -q -x -k "tf32bf16" --tb=short 2>&1 | tail -n 12 > gpurun_out/mixed_ops.log; cat gpurun_out/mixed_ops.log
for b in 7 64; do
  for g in tf32x3 tf32bf16; do
    timeout 200 python bench.py --gemm $g --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-device-dataset > gpurun_out/bench_${g}_b$b.json 2> gpurun_out/bench_mixed.err
    python - $g $b <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench_{sys.argv[1]}_b{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "batch", sys.argv[2], "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "roof", round(d.get("roofline",{}).get("frac",0),4))
except Exception as e: print("bench unreadable", e)
PY
  done
done
M2D_GEMM=tf32bf16 timeout 200 python -m pytest tests/test_parity_gpu.py -q 2>&1 | grep -E "^E  .*(Error|assert)|^FAILED|passed|failed" | tail -n 12 > gpurun_out/parity_mixed.log; cat gpurun_out/parity_mixed.log
