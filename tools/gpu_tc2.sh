#!/bin/bash
set -u
mkdir -p gpurun_out
{
for mode in tf32 tf32x3; do timeout 120 python tools/tc_probe.py wgrad $mode 2>&1 | tail -40; echo "rc=$? (wgrad $mode)"; done
} > gpurun_out/tc_probe2.log 2>&1
timeout 900 python -m pytest tests/test_ops_gpu.py -q 2>&1 | tail -40 > gpurun_out/pytest_ops2.log
tail -40 gpurun_out/tc_probe2.log; tail -15 gpurun_out/pytest_ops2.log
