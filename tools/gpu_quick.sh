#!/bin/bash
# fused-trainer parity tests + one bench line per batch size
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "fused or additivity" 2>&1 | tail -n 4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64.json 2>> gpurun_out/bench.err
tail -n 3 gpurun_out/bench.err
for f in bench bench_b64; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches/step", d["gpu_launches_per_step"], "roof", d.get("roofline",{}).get("kernel"), round(d.get("roofline",{}).get("frac",0),4))
    for k,v in sorted(d.get("kernel_families",{}).items(), key=lambda kv:-kv[1]["ms"]):
        print("   %-14s n=%5d ms=%8.3f tflops=%7.2f gbs=%8.1f share=%.3f"%(k,v["launches"],v["ms"],v["tflops"],v["gbs"],v["share_of_eager_step"]))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
