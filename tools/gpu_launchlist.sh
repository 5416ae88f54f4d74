#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 4200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
tail -n 2 gpurun_out/ncu_bench.log
