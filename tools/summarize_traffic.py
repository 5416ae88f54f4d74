#!/usr/bin/env python
"""ncu per-launch DRAM traffic of the tensor-core GEMM families -> profiles/*.json (read by bench.py's
roofline.traffic), and a markdown table of one `ncu --set full` report per distinct grid.
    python tools/summarize_traffic.py traffic gpurun_out/tc_traffic.csv profiles/r01_tc_traffic.json
    python tools/summarize_traffic.py full gpurun_out/prof_halo_r01.ncu-rep profiles/r01_ncu_full_rowconv_halo.md "<title>"
"""
import collections
import csv
import json
import subprocess
import sys

FAMILY = {"rowconv_halo_kernel": "rowconv", "rowconv_tc_kernel": "rowconv", "wgrad_tc_kernel": "wgrad"}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def traffic(src, dst):
    lines = open(src).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    per = collections.defaultdict(dict)
    for r in csv.DictReader(lines[start:]):
        v = num(r["Metric Value"])
        if v is None:
            continue
        u = r["Metric Unit"]
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0}.get(u, 1.0)
        per[r["ID"]][r["Metric Name"]] = v
        per[r["ID"]]["kernel"] = r["Kernel Name"]
    fam = collections.defaultdict(lambda: dict(n=0, rd=0.0, wr=0.0, us=0.0, tp=0.0, by_kernel=collections.Counter()))
    for d in per.values():
        k = next((f for f in FAMILY if f in d["kernel"]), None)
        if k is None:
            continue
        f = fam[FAMILY[k]]
        f["n"] += 1
        f["rd"] += d.get("dram__bytes_read.sum", 0.0)
        f["wr"] += d.get("dram__bytes_write.sum", 0.0)
        t = d.get("gpu__time_duration.sum", 0.0)
        f["us"] += t
        f["tp"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        f["by_kernel"][k] += 1
    out = {}
    for name, f in fam.items():
        out[name] = {"launches": f["n"], "kernels": dict(f["by_kernel"]),
                     "dram_read_bytes_per_launch": f["rd"] / f["n"], "dram_write_bytes_per_launch": f["wr"] / f["n"],
                     "traffic_bytes_per_launch": (f["rd"] + f["wr"]) / f["n"],
                     "avg_us_per_launch_under_ncu": f["us"] / f["n"],
                     "tensor_pipe_active_pct_time_weighted": f["tp"] / max(f["us"], 1e-9)}
    out["source"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,"
                     "sm__pipe_tensor_cycles_active... -k rowconv_halo|rowconv_tc|wgrad_tc -s 1800 -c 900, bench.py --steps 1 "
                     "--warmup 1 --no-graphs (batch 7, default encoder, gemm tf32x3); tools/gpu_traffic.sh + "
                     "tools/summarize_traffic.py")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed.sum.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def full(rep, dst, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen, cols = set(), []
    for d in data:
        key = (d[idx["Kernel Name"]].split("(")[0], d[idx["Grid Size"]])
        if key not in seen:
            seen.add(key)
            cols.append(d)
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nsource: `{rep}` (`ncu --set full --clock-control none --import-source on`); one column per "
                "distinct (kernel, grid)\n\n")
        f.write("| metric | unit | " + " | ".join(f"{d[idx['Kernel Name']].split('(')[0].replace('void m2d::', '')[:24]} "
                                                  f"{d[idx['Grid Size']]}" for d in cols) + " |\n")
        f.write("|---|---|" + "---|" * len(cols) + "\n")
        for m in WANT:
            if m in idx:
                f.write(f"| `{m}` | {units[idx[m]]} | " + " | ".join(d[idx[m]] for d in cols) + " |\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
