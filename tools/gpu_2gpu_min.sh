#!/bin/bash
# 2 GPUs, final-tree check: the data-parallel step against the sharded oracle (both collectives) + the batch-7 bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -q --tb=short 2>&1 | grep -vE "^$|Warning|warn|Consider" | tail -n 6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline --no-device-dataset --no-throughput-regime > gpurun_out/bench_2gpu_nvl.json 2> gpurun_out/bench_2gpu_nvl.err
echo "bench rc=$?"; grep -vE "OMP_NUM|\*\*\*\*|^$" gpurun_out/bench_2gpu_nvl.err | tail -n 3 | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-roofline --no-device-dataset > gpurun_out/bench_1gpu.json 2>gpurun_out/bench_1gpu.err
python - <<'PY'
import json,glob
for f in ("gpurun_out/bench_1gpu.json", "gpurun_out/bench_2gpu_nvl.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-36s n=%d value %8.3f ms/step %7.3f e2e %8.3f seq/s %8.1f | %s"%(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["sequences_per_s"], (d["config"].get("collective") or "")[:60]))
    except Exception as e: print(f, "unreadable", e)
PY
