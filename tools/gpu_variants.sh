#!/bin/bash
# single-pass TF32 mode (parity suite + bench) and the other two audio encoders (BASELINE configs[4]: wavegan; unet)
set -u
mkdir -p gpurun_out
M2D_GEMM=tf32 timeout 600 python -m pytest tests/test_parity_gpu.py -q 2>&1 | grep -E "^E  .*(Error|assert)|^FAILED|passed|failed" > gpurun_out/parity_tf32.log
tail -n 30 gpurun_out/parity_tf32.log
timeout 300 python bench.py --gemm tf32 --steps 20 --warmup 3 --no-cpu-baseline --no-device-dataset > gpurun_out/bench_tf32.json 2> gpurun_out/bench_var.err
timeout 300 python bench.py --enc wavegan --steps 20 --warmup 3 --no-cpu-baseline --no-device-dataset > gpurun_out/bench_wavegan.json 2>> gpurun_out/bench_var.err
timeout 300 python bench.py --enc unet --steps 10 --warmup 3 --no-cpu-baseline --no-device-dataset > gpurun_out/bench_unet.json 2>> gpurun_out/bench_var.err
tail -n 5 gpurun_out/bench_var.err
for f in bench_tf32 bench_wavegan bench_unet; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches/step", d["gpu_launches_per_step"], "roof", d.get("roofline",{}).get("kernel"), round(d.get("roofline",{}).get("frac",0),4))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
