#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k gru 2>&1 | tail -n 15 > gpurun_out/gru2.log; cat gpurun_out/gru2.log
for mode in fp32 tf32x3; do
  M2D_GEMM=$mode timeout 900 python -m pytest tests/test_parity_gpu.py -q 2>&1 | grep -E "^E  .*Error|^FAILED|passed|failed" > gpurun_out/parity_$mode.log
  echo "== $mode"; tail -n 40 gpurun_out/parity_$mode.log
done
