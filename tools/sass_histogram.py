#!/usr/bin/env python
"""SASS opcode histogram per kernel of libm2d_b200.so (cuobjdump -sass): which kernels carry the Blackwell-native
instructions (tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, tcgen05.commit = UTCBAR, TMA = UTMALDG / UBLKCP,
multimem.ld_reduce = LDGMC, st.async / DSMEM, clusters) and which are CUDA-core kernels.
    python tools/sass_histogram.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "music2dance_b200", "libm2d_b200.so")
KEYS = [("UTCHMMA", r"^UTCHMMA"), ("UTC*MMA other", r"^UTC(?!HMMA|BAR)[A-Z]*MMA"), ("LDTM", r"^LDTM"), ("STTM", r"^STTM"),
        ("UTCBAR", r"^UTCBAR"), ("UTMALDG", r"^UTMALDG"), ("UBLKCP", r"^UBLKCP"), ("LDGMC (multimem)", r"^LDGMC"),
        ("ATOMG.CAS.SYS", r"^ATOMG.*CAS.*SYS"), ("SYNCS (mbarrier)", r"^SYNCS"), ("UCGABAR/CGA", r"^(UCGABAR|CGABAR|CGAERRBAR)"),
        ("STAS (st.async)", r"^STAS"), ("HMMA (legacy mma)", r"^HMMA"), ("FFMA", r"^FFMA"), ("LDGSTS", r"^LDGSTS"),
        ("BAR.SYNC", r"^BAR")]
txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
kern, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["_total"] += 1
        for name, pat in KEYS:
            if re.match(pat, op):
                kern[cur][name] += 1


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = re.sub(r"\(.*", "", r).replace("void m2d::", "").replace("m2d::", "")
    return r[:70]


agg = collections.OrderedDict()
for n, c in kern.items():
    d = demangle(n)
    base = re.sub(r"<.*", "", d)
    a = agg.setdefault(base, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
out = ["# round 2: SASS opcode histogram of `music2dance_b200/libm2d_b200.so`", "",
       f"`cuobjdump -sass` of the in-tree library (cubins: {', '.join(arch)} only), opcode counts summed over the template "
       "instances of each kernel (`tools/sass_histogram.py`).  tcgen05.mma -> `UTCHMMA`, tcgen05.ld -> `LDTM`, "
       "tcgen05.commit -> `UTCBAR`, TMA tensor / bulk copies -> `UTMALDG` / `UBLKCP`, multimem.ld_reduce -> `LDGMC`, "
       "mbarrier -> `SYNCS`, st.async -> `STAS`.  No `HMMA` (legacy mma.sync) anywhere.", "",
       "| kernel | instances | SASS instr. | " + " | ".join(k for k, _ in KEYS) + " |",
       "|---|---:|---:|" + "---:|" * len(KEYS)]
tot = collections.Counter()
for base, (ni, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]["_total"]):
    out.append(f"| `{base}` | {ni} | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k, _ in KEYS) + " |")
    tot.update(c)
out.append(f"| **all {len(kern)} kernels** | | {tot['_total']} | " + " | ".join(str(tot[k]) for k, _ in KEYS) + " |")
dst = sys.argv[1] if len(sys.argv) > 1 else None
s = "\n".join(out) + "\n"
if dst:
    open(dst, "w").write(s)
print(s)
