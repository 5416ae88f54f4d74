#!/bin/bash
# tcgen05 bring-up: probes in separate processes, then the kernel-level parity tests
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
{
for fam in fwd wgrad; do for mode in tf32 tf32x3; do
  timeout 120 python tools/tc_probe.py $fam $mode 2>&1 | tail -40; echo "rc=$? ($fam $mode)"
done; done
for mode in fp32 tf32 tf32x3; do timeout 120 python tools/tc_probe.py timing $mode 2>&1 | tail -12; done
} > gpurun_out/tc_probe.log 2>&1
timeout 900 python -m pytest tests/test_ops_gpu.py -q 2>&1 | tail -40 > gpurun_out/pytest_ops.log
tail -60 gpurun_out/tc_probe.log; tail -15 gpurun_out/pytest_ops.log
