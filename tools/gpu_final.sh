#!/bin/bash
# Round-end state check sized for a small GPU budget: all GPU tests, the default bench line, one large-batch line,
# then (if time remains) the ncu launch list of the same bench command.
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  .*Error|^FAILED|passed|failed" > gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
timeout 120 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 80 python bench.py --batch 512 --steps 3 --warmup 3 --no-cpu-baseline --no-device-dataset > gpurun_out/bench_b512.json 2>> gpurun_out/bench.err; echo "b512 rc=$?" >> gpurun_out/bench.err
tail -n 4 gpurun_out/bench.err | cut -c1-300
for f in bench bench_b512; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches/step", d["gpu_launches_per_step"], "roof", d.get("roofline",{}).get("kernel"), round(d.get("roofline",{}).get("frac",0),4), "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 4200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline --no-device-dataset > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
