#!/bin/bash
# round 2 evidence: (1) ncu launch list of the bench command, (2) per-family DRAM traffic of exactly one train step,
# (3) ncu --set full of the kernels the roofline table names, (4) SASS opcode histogram per kernel
set -u
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
for b in 7 64; do
  for fam in rowconv wgrad; do
    timeout 900 ncu --profile-from-start off --nvtx --nvtx-include "$fam/" --metrics $M --clock-control none --csv --log-file gpurun_out/traffic_${fam}_b$b.csv python tools/traffic_step.py $b > gpurun_out/traffic_${fam}_b$b.log 2>&1
    python tools/summarize_traffic2.py $fam $b default tf32x3 gpurun_out/traffic_${fam}_b$b.csv gpurun_out/traffic_counts_b$b.json gpurun_out/r02_traffic.json | tail -n 12
  done
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 3600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline --no-eager-gpu --no-throughput-regime --no-device-dataset > gpurun_out/ncu_bench.log 2>&1
tail -n 2 gpurun_out/ncu_bench.log | cut -c1-300
for k in wgrad_tc_kernel gru_fwd2_kernel gru_bwd2_kernel conv_c1_fwd_kernel conv_c1_wgrad_kernel conv_c1_dgrad4_kernel adam_pack_kernel rowconv_halo_persist_kernel rowconv_halo_kernel skinny_gemm_kernel colsum_batch_kernel; do
  # the reports stay on the box (64 MiB merge limit); their raw pages come back as CSV
  M2D_OVERLAP=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 3 -f -o /tmp/full_$k python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline --no-eager-gpu --no-throughput-regime --no-device-dataset > gpurun_out/ncu_full_$k.log 2>&1
  ncu -i /tmp/full_$k.ncu-rep --page raw --csv > gpurun_out/full_$k.csv 2>/dev/null
  ls -la /tmp/full_$k.ncu-rep gpurun_out/full_$k.csv 2>/dev/null | awk '{print $5, $9}'
done
rm -f gpurun_out/ncu_full_*.log
