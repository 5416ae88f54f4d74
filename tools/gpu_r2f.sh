#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "persistent" --tb=short 2>&1 | grep -E "Error|assert|passed|failed|skipped" | cut -c1-400 | tail -n 12
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  .*|^FAILED|passed|failed|Error" | head -n 20
for b in 7 64; do
timeout 600 python bench.py --batch $b --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_b$b.err
M2D_HALO_PERSIST=0 timeout 600 python bench.py --batch $b --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset > gpurun_out/bench_b${b}_nopersist.json 2> gpurun_out/bench_b$b.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_b*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches", d.get("gpu_launches_per_step"), "roof", round(d.get("roofline",{}).get("frac",0),4))
        for k,v in sorted(d.get("kernel_families",{}).items(), key=lambda kv:-kv[1]["ms"])[:3]:
            print("   %-14s n=%5d ms=%8.3f tflops=%7.2f share=%.3f"%(k,v["launches"],v["ms"],v["tflops"],v["share_of_eager_step"]))
    except Exception as e: print(f, "unreadable", e)
PY
