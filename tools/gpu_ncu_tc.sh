#!/bin/bash
# ncu --set full of the tensor-core rowconv kernel on a big shape (probe), source-level
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowconv_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_rowconv_tc python tools/tc_probe.py timing tf32x3 > gpurun_out/ncu_tc.log 2>&1
tail -n 5 gpurun_out/ncu_tc.log
ls -la gpurun_out/*.ncu-rep
