#!/bin/bash
# DRAM traffic + duration of every tensor-core GEMM launch of ~1 train step (roofline.traffic), and an
# ncu --set full capture of 40 row-convolution launches inside a train step (single stream)
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"rowconv_halo_kernel|rowconv_tc_kernel|wgrad_tc_kernel" -s 1800 -c 900 --csv --log-file gpurun_out/tc_traffic.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline > gpurun_out/ncu_traffic.log 2>&1
tail -n 1 gpurun_out/ncu_traffic.log | cut -c1-200
M2D_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowconv_halo_kernel -s 1500 -c 40 -f -o gpurun_out/prof_rowconv_halo_r01 python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
