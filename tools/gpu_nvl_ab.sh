#!/bin/bash
# 2 GPUs: the peer-memory all-reduce alone (tools/nvl_check.py: exactness vs NCCL + time) and the batch-7 bench line, for
# one vector vs four in flight per thread and 32 vs 64 blocks
set -u
mkdir -p gpurun_out
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for u in 0 1; do for b in 32 64; do
  echo "== unroll $u blocks $b"
  M2D_NVL_UNROLL=$u M2D_NVL_BLOCKS=$b tr 29541 tools/nvl_check.py 2>&1 | grep -E "n = |rank .: (OK|FAILED)|MISMATCH|Error" | head -8
done; done
for cfg in "0 32" "1 32" "1 64"; do set -- $cfg
  M2D_NVL_UNROLL=$1 M2D_NVL_BLOCKS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 15 --warmup 3 --no-roofline --no-device-dataset --no-throughput-regime > gpurun_out/bench_2gpu_u$1_b$2.json 2> gpurun_out/bench_2gpu.err
  python - $1 $2 <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench_2gpu_u{sys.argv[1]}_b{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(f"unroll {sys.argv[1]} blocks {sys.argv[2]}: value {d['value']:.3f} ms/step {d['ms_per_step']:.3f}")
except Exception as e: print("unreadable", e, open("gpurun_out/bench_2gpu.err").read()[-600:])
PY
done
