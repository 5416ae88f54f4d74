#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "fused_trainer or resynced or additivity" 2>&1 | tail -n 2
timeout 300 python tools/step_timeline.py 7 gpurun_out/timeline_b7.json > gpurun_out/timeline_b7.txt 2>&1; grep -E "step |pose_wg_end|aud_wg_l6|bwd_end|adam_pack|aud_tan_l6|start-to-start" gpurun_out/timeline_b7.txt | head -n 8
timeout 600 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset --no-roofline > gpurun_out/bench_b7.json 2> gpurun_out/bench_b7.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_b7.json").read().strip().splitlines()[-1])
print("value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches", d.get("gpu_launches_per_step"))
PY
