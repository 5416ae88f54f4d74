#!/bin/bash
# 4 GPUs, final tree: the all-reduce kernel against NCCL (tools/nvl_check.py) and the batch-7 bench line
set -u
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 tools/nvl_check.py 2>&1 | grep -E "n = |rank .: (OK|FAILED)|MISMATCH|Error" | head -8
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 10 --warmup 3 --no-roofline --no-device-dataset --no-throughput-regime > gpurun_out/bench_4gpu_nvl.json 2> gpurun_out/bench_4gpu_nvl.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_4gpu_nvl.json").read().strip().splitlines()[-1])
    print("n=%d value %.3f ms/step %.3f e2e %.3f seq/s %.1f"%(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["sequences_per_s"]))
except Exception as e: print("unreadable", e, open("gpurun_out/bench_4gpu_nvl.err").read()[-600:])
PY
