#!/bin/bash
# phase2 conditional bring-up: its GPU tests first (full traceback), then the whole GPU suite (summary only)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_phase2_cond.py -m gpu -q -x --tb=short 2>&1 | tail -n 40 > gpurun_out/cond.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  .*Error|^FAILED|passed|failed" > gpurun_out/pytest_gpu.log
cat gpurun_out/cond.log; tail -n 15 gpurun_out/pytest_gpu.log
