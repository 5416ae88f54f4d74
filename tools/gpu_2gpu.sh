#!/bin/bash
# 2 GPUs: new collective schedule (early bucket next to the l5/l6 weight gradients, plain + arena in one launch)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py tests/test_parity_gpu.py -q --tb=short -k "two_gpu or fused_trainer or dropin_step" 2>&1 | grep -vE "^$|Warning|warn|Consider|out = dict" | tail -n 8
timeout 300 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-roofline --no-device-dataset > gpurun_out/bench_1gpu.json 2>gpurun_out/bench_1gpu.err
for c in nvl nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --collective $c --no-roofline --no-device-dataset > gpurun_out/bench_2gpu_$c.json 2> gpurun_out/bench_2gpu_$c.err
  echo "bench $c rc=$?"; grep -vE "OMP_NUM|\*\*\*\*|^$" gpurun_out/bench_2gpu_$c.err | tail -n 3 | cut -c1-300
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --batch 64 --steps 5 --warmup 3 --no-roofline --no-device-dataset > gpurun_out/bench_2gpu_b64.json 2> gpurun_out/bench_2gpu_b64.err
timeout 300 python bench.py --batch 64 --steps 5 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-roofline --no-device-dataset > gpurun_out/bench_1gpu_b64.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*gpu*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s value %8.3f ms/step %7.3f e2e %8.3f seq/s %8.1f | %s"%(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["sequences_per_s"], (d["config"].get("collective") or "")[:60]))
    except Exception as e: print(f, "unreadable", e)
PY
