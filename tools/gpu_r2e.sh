#!/bin/bash
# 2 GPUs: peer-memory all-reduce kernel — parity against the sharded oracle, then bench lines for both collectives
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 200 python -m pytest tests/test_ops_gpu.py -q -k "adam_pack" --tb=short 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-600
timeout 900 python -m pytest tests/test_dp_gpu.py -q -s --tb=short 2>&1 | grep -vE "^$|Warning|warn" | tail -n 25
for c in nvl nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --collective $c --no-roofline --no-device-dataset > gpurun_out/bench_2gpu_$c.json 2> gpurun_out/bench_2gpu_$c.err
  echo "bench $c rc=$?"; tail -n 4 gpurun_out/bench_2gpu_$c.err | cut -c1-400
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-roofline --no-device-dataset > gpurun_out/bench_1gpu.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*gpu*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), d["config"].get("collective"), d["config"].get("graph_structure"))
    except Exception as e: print(f, "unreadable", e)
PY
