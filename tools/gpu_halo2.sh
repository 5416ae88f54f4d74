#!/bin/bash
set -u
mkdir -p gpurun_out
for na in 3 4; do echo "=== timing halo on, NA=$na"; M2D_HALO_NA=$na timeout 300 python tools/halo_probe.py timing tf32x3 2>&1 | grep -v dgrad | tail -20; done
echo "=== tf32 single NA=3";  M2D_HALO_NA=3 timeout 300 python tools/halo_probe.py timing tf32 2>&1 | grep -v dgrad | tail -20
