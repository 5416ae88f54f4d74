#!/usr/bin/env python
"""`ncu -i <rep> --page raw --csv` dumps of `--set full` captures (one per kernel, tools/gpu_profiles.sh) -> ONE
markdown table, one column per (kernel, grid) — the metrics B200_PROFILING.md names.
    python tools/summarize_full_csv.py profiles/r02_ncu_full.md gpurun_out/full_*.csv"""
import csv
import os
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed.sum.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_selected", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
dst, files = sys.argv[1], sys.argv[2:]
cols, units_of = [], {}
for f in sorted(files):
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0].replace("void m2d::", "").replace("m2d::", "")
        key = (name, d[idx["Grid Size"]], d[idx["Block Size"]])
        if key in seen:
            continue
        seen.add(key)
        cols.append((key, {m: d[idx[m]] for m in WANT if m in idx}))
        for m in WANT:
            if m in idx:
                units_of[m] = units[idx[m]]
with open(dst, "w") as out:
    out.write("# round 2: `ncu --set full` per kernel (batch 7, default encoder, tf32x3, single stream, eager)\n\n"
              "source: `tools/gpu_profiles.sh` — one `ncu --set full --clock-control none --import-source on -k regex:<kernel> "
              "-s 20 -c 3` capture per kernel over `bench.py --steps 1 --warmup 1 --no-graphs`; the reports stay on the GPU "
              "box, their raw pages come back as CSV (`ncu -i ... --page raw --csv`); one column per distinct (kernel, grid).\n\n")
    out.write("| metric | unit | " + " | ".join(f"`{k[0][:34]}` {k[1]} x {k[2]}" for k, _ in cols) + " |\n")
    out.write("|---|---|" + "---|" * len(cols) + "\n")
    for m in WANT:
        if m in units_of:
            out.write(f"| `{m}` | {units_of[m]} | " + " | ".join(v.get(m, "") for _, v in cols) + " |\n")
print(open(dst).read()[:3000])
