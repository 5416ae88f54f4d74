#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  .*|^FAILED|passed|failed|Error" | head -n 30
timeout 200 python tools/adam_pack_probe.py 2>&1 | grep -E "adam_pack|items"
M2D_AP_ROWS=16 timeout 200 python tools/adam_pack_probe.py 2>&1 | grep -E "adam_pack|items"
M2D_AP_ROWS=16 timeout 200 python -m pytest tests/test_ops_gpu.py -q -k adam_pack 2>&1 | tail -n 2
timeout 600 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset > gpurun_out/bench_b7.json 2> gpurun_out/bench_b7.err
M2D_AP_ROWS=16 timeout 600 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset > gpurun_out/bench_b7_ap16.json 2> gpurun_out/bench_b7.err
python - <<'PY'
import json,glob
for f in ["gpurun_out/bench_b7.json","gpurun_out/bench_b7_ap16.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s value %8.3f ms/step %8.2f e2e %8.3f launches %5s"%(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches_per_step")))
    except Exception as e: print(f, "unreadable", e)
PY
bash tools/gpu_sanitize.sh
