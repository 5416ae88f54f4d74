#!/bin/bash
# round 2, call b: fused Adam + re-layout kernel — bit-exactness test, whole GPU suite, bench, timeline
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "adam_pack" --tb=short 2>&1 | tail -n 15
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  .*|^FAILED|passed|failed|Error" | head -n 30
timeout 300 python tools/step_timeline.py 7 gpurun_out/timeline_b7.json > gpurun_out/timeline_b7.txt 2>&1; head -n 45 gpurun_out/timeline_b7.txt; tail -n 3 gpurun_out/timeline_b7.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline > gpurun_out/bench_b7.json 2> gpurun_out/bench_b7.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_b7.err
python - <<'PY'
import json,glob
for f in ["gpurun_out/bench_b7.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches", d.get("gpu_launches_per_step"), "roof", d.get("roofline",{}).get("kernel"), round(d.get("roofline",{}).get("frac",0),4))
        if "throughput_regime" in d: print("    regime", {k:d["throughput_regime"].get(k) for k in ("batch_per_gpu","value","ms_per_step","sequences_per_s")}, (d["throughput_regime"].get("roofline") or {}).get("frac"))
    except Exception as e: print(f, "unreadable", e)
PY
