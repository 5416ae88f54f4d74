#!/usr/bin/env python
"""ncu CSV (metrics per launch, NVTX-filtered to one kernel family of ONE train step, tools/traffic_step.py) ->
entry of profiles/r02_traffic.json that bench.py's roofline.traffic reads.
    python tools/summarize_traffic2.py <family> <batch> <enc> <gemm> <ncu.csv> <counts.json> <out.json>
"""
import collections
import csv
import json
import os
import sys

fam, batch, enc, gemm, src, counts, dst = sys.argv[1:8]
lines = open(src).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
per = collections.defaultdict(dict)
for r in csv.DictReader(lines[start:]):
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0}.get(r["Metric Unit"], 1.0)
    per[r["ID"]][r["Metric Name"]] = v
    per[r["ID"]]["kernel"] = r["Kernel Name"].split("(")[0].replace("void m2d::", "")
calls = json.load(open(counts))["calls_per_step"][fam]
rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in per.values())
wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in per.values())
us = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
tp = sum(d.get("gpu__time_duration.sum", 0.0) * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
         for d in per.values())
kern = collections.Counter(d["kernel"].split("<")[0] for d in per.values())
out = json.load(open(dst)) if os.path.exists(dst) else {}
out[f"{fam}:b{batch}:{enc}:{gemm}"] = {
    "launches_per_step": calls, "kernels_per_step": len(per), "kernels": dict(kern),
    "dram_read_bytes_per_launch": rd / calls, "dram_write_bytes_per_launch": wr / calls,
    "traffic_bytes_per_launch": (rd + wr) / calls, "avg_us_per_launch_under_ncu": us / calls,
    "tensor_pipe_active_pct_time_weighted": tp / max(us, 1e-9),
    "source": "ncu --profile-from-start off --nvtx --nvtx-include '%s/' --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
              "gpu__time_duration.sum,sm__pipe_tensor_cycles_active... over ONE eager single-stream train step "
              "(tools/traffic_step.py, tools/gpu_profiles.sh)" % fam}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out[f"{fam}:b{batch}:{enc}:{gemm}"], indent=1))
