#!/bin/bash
# bring-up of the halo kernel: correctness probe, timings, kernel-level tests
set -u
mkdir -p gpurun_out
echo "=== check tf32x3"; timeout 200 python tools/halo_probe.py check tf32x3 2>&1 | tail -24 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12}' | cut -c1-80
echo "=== timing halo on"; timeout 300 python tools/halo_probe.py timing tf32x3 2>&1 | tail -20
echo "=== timing halo off"; M2D_HALO=0 timeout 300 python tools/halo_probe.py timing tf32x3 2>&1 | tail -20
echo "=== timing halo on tf32"; timeout 300 python tools/halo_probe.py timing tf32 2>&1 | grep -v dgrad | tail -20
timeout 900 python -m pytest tests/test_ops_gpu.py -q -x 2>&1 | tail -n 5
