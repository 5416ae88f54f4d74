#!/bin/bash
# compute-sanitizer over the kernel-level tests (SURVEY §5 "race detection"): memcheck, racecheck (shared-memory hazards of
# the hand-rolled mbarrier / TMA / tcgen05 pipelines), synccheck.  One arithmetic mode (the default 3xTF32) to bound time.
set -u
mkdir -p gpurun_out
SEL='tf32x3 and (conv_fwd or merged or epilogue or full_length or linear or windowed or gru or batchnorm or elementwise or adam or persistent)'
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_ops_gpu.py -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_$tool.log | tail -n 6
done
