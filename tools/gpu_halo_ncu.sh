#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowconv_halo -c 8 -f -o gpurun_out/prof_halo python tools/halo_probe.py prof tf32x3 > gpurun_out/ncu_halo.log 2>&1
tail -3 gpurun_out/ncu_halo.log; ls -la gpurun_out/prof_halo.ncu-rep
