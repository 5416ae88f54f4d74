#!/bin/bash
# 8 GPUs of one box: the data-parallel step with the peer-memory all-reduce at N = 8 and 4 (and NCCL at 8 for comparison)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8 | tr '\n' ';'; echo
run() { local n=$1 c=$2 b=$3 name=$4
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --batch $b --steps 10 --warmup 3 --collective $c --no-roofline --no-device-dataset > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name rc=$?"; grep -vE "OMP_NUM|\*\*\*\*|^$|Warning|warn" gpurun_out/bench_$name.err | tail -n 2 | cut -c1-300; }
run 8 nvl 7 8gpu_nvl
run 8 nccl 7 8gpu_nccl
run 4 nvl 7 4gpu_nvl
run 8 nvl 64 8gpu_nvl_b64
timeout 200 python bench.py --steps 10 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-roofline --no-device-dataset > gpurun_out/bench_1gpu_ref.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_[1248]gpu_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s n=%d value %8.3f ms/step %7.3f e2e %8.3f seq/s %8.1f | %s"%(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["sequences_per_s"], (d["config"].get("collective") or "")[:50]))
    except Exception as e: print(f, "unreadable", e)
PY
