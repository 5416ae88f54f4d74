#!/bin/bash
# ncu source pages (per-SASS-instruction stall samples) of selected kernels inside an eager train step
set -u
mkdir -p gpurun_out
for k in rowconv_halo_kernel wgrad_tc_kernel adam_pack_kernel rowconv_tc_kernel; do
  M2D_OVERLAP=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 6 -f -o /tmp/src_$k python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline --no-eager-gpu --no-throughput-regime --no-device-dataset > /dev/null 2>&1
  ncu -i /tmp/src_$k.ncu-rep --page source --csv > gpurun_out/src_$k.csv 2>/dev/null
  ls -la gpurun_out/src_$k.csv | awk '{print $5, $9}'
done
