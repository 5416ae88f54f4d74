#!/bin/bash
# Evidence refresh on the final tree (batch 7): DRAM traffic of the rowconv / wgrad launches of ONE train step (feeds
# roofline.traffic), the ncu launch list of the bench command, the in-graph timeline.  Results -> gpurun_out/.
set -u
mkdir -p gpurun_out
cp profiles/r02_traffic.json gpurun_out/r02_traffic.json
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
for fam in rowconv wgrad; do
  timeout 900 ncu --profile-from-start off --nvtx --nvtx-include "$fam/" --metrics $M --clock-control none --csv --log-file gpurun_out/traffic_${fam}_b7.csv python tools/traffic_step.py 7 > gpurun_out/traffic_${fam}_b7.log 2>&1
  python tools/summarize_traffic2.py $fam 7 default tf32x3 gpurun_out/traffic_${fam}_b7.csv gpurun_out/traffic_counts_b7.json gpurun_out/r02_traffic.json | tail -n 12
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 2600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline --no-eager-gpu --no-throughput-regime --no-device-dataset > gpurun_out/ncu_bench.log 2>&1
tail -n 2 gpurun_out/ncu_bench.log | cut -c1-300
python tools/step_timeline.py 7 gpurun_out/timeline_b7.json > gpurun_out/timeline_b7.txt 2>&1; tail -n 3 gpurun_out/timeline_b7.txt
