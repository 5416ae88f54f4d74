#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; local b=$1; shift
  env "$@" timeout 600 python bench.py --batch $b --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset --no-roofline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; }
run b7_prio 7 M2D_PRIO=1
run b7_noprio 7 M2D_PRIO=0
run b7_prio_again 7 M2D_PRIO=1
run b64_prio 64 M2D_PRIO=1
run b64_noprio 64 M2D_PRIO=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_b*prio*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s value %8.3f ms/step %8.2f e2e %8.3f"%(f, d["value"], d["ms_per_step"], d["e2e"]["value"]))
    except Exception as e: print(f, "unreadable", e, open(f.replace('.json','.err')).read()[-600:])
PY
timeout 300 python tools/step_timeline.py 7 gpurun_out/timeline_b7.json > gpurun_out/timeline_b7.txt 2>&1; head -n 45 gpurun_out/timeline_b7.txt | tail -n 43; grep -E "start-to-start|gu:|gb:" gpurun_out/timeline_b7.txt | head -n 9
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -n 3
