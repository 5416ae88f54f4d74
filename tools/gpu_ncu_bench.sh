#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowconv_tc_kernel -s 120 -c 24 -f -o gpurun_out/prof_rowconv_b64 python bench.py --batch 64 --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline > gpurun_out/ncu_b64.log 2>&1
tail -n 2 gpurun_out/ncu_b64.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
