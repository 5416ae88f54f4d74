#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "adam_pack" --tb=short 2>&1 | tail -n 15
timeout 200 python tools/adam_pack_probe.py 2>&1 | tee gpurun_out/adam_pack_probe.txt
timeout 300 python tools/step_timeline.py 7 gpurun_out/timeline_b7.json > gpurun_out/timeline_b7.txt 2>&1; grep -E "step |bwd_end|adam_pack|start-to-start|gu:|gb:" gpurun_out/timeline_b7.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime > gpurun_out/bench_b7.json 2> gpurun_out/bench_b7.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_b7.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_b7.json").read().strip().splitlines()[-1])
print("value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches", d.get("gpu_launches_per_step"))
PY
