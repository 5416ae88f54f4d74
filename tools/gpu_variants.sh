#!/bin/bash
# other encoders (BASELINE configs[4] = wavegan; unet) at batch 7 on one GPU, long-sequence generator inference
set -u
mkdir -p gpurun_out
for e in wavegan unet; do
  timeout 600 python bench.py --enc $e --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset > gpurun_out/bench_b7_$e.json 2> gpurun_out/bench_$e.err
done
timeout 300 python tools/long_sequences.py > gpurun_out/long_sequences.jsonl 2> gpurun_out/long.err; tail -n 3 gpurun_out/long.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_b7_*net.json"))+sorted(glob.glob("gpurun_out/bench_b7_wavegan.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s value %8.3f ms/step %7.3f e2e %8.3f launches %d roof %.4f"%(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"], d["roofline"]["frac"]))
    except Exception as e: print(f, "unreadable", e)
print(open("gpurun_out/long_sequences.jsonl").read()[-1500:])
PY
