#!/bin/bash
# TF32 + BF16 cross-term mode (branch next/tf32-bf16-split): kernel-level tests, a bench line, then the parity suite
set -u
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_ops_gpu.py -q -x -k "tf32bf16" --tb=short 2>&1 | tail -n 12 > gpurun_out/mixed_ops.log; cat gpurun_out/mixed_ops.log
timeout 50 python bench.py --gemm tf32bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --no-device-dataset > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_mixed.json").read().strip().splitlines()[-1])
    print("mixed value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), d["last_step_logs"])
except Exception as e: print("bench unreadable", e)
PY
M2D_GEMM=tf32bf16 timeout 60 python -m pytest tests/test_parity_gpu.py -q 2>&1 | grep -E "^E  .*(Error|assert)|^FAILED|passed|failed" | tail -n 12 > gpurun_out/parity_mixed.log; cat gpurun_out/parity_mixed.log
