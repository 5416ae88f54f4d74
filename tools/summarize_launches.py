#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table:
    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches.md "<title>"
ncu times are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys


def main(src, dst, title):
    with open(src) as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot, n = 0.0, 0
    for r in csv.DictReader(lines[start:]):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        mine = "m2d::" in name
        key = (name, r["Grid Size"], r["Block Size"], mine)
        agg[key][0] += 1
        agg[key][1] += v
        tot += v
        n += 1
    fam = collections.defaultdict(lambda: [0, 0.0])
    for (name, g, b, mine), (c, us) in agg.items():
        k = re.sub(r"<.*", "", name) if mine else "(torch) " + re.sub(r"<.*", "", name)[:60]
        fam[k][0] += c
        fam[k][1] += us
    with open(dst, "w") as o:
        o.write(f"# {title}\n\nsource: `{src}` — {n} launches, {tot / 1e3:.2f} ms of device time "
                "(ncu per-launch times: cold cache, serialised — shares, not absolutes)\n\n")
        o.write("## by kernel\n\n| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (c, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            o.write(f"| `{k}` | {c} | {us / 1e3:.3f} | {us / tot:.3f} | {us / c:.1f} |\n")
        o.write("\n## top 40 (kernel, grid, block)\n\n| kernel | grid | block | launches | total ms | share | avg us |\n"
                "|---|---|---|---:|---:|---:|---:|\n")
        for (name, g, b, mine), (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            o.write(f"| `{name}` | {g} | {b} | {c} | {us / 1e3:.3f} | {us / tot:.3f} | {us / c:.1f} |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu launch list")
