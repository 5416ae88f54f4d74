#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "persistent" --tb=short 2>&1 | grep -E "Error|assert|passed|failed|skipped" | cut -c1-400 | tail -n 8
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  .*|^FAILED|passed|failed|Error" | head -n 20
run() { # name, env..., batch
  local name=$1; shift; local b=$1; shift
  env "$@" timeout 600 python bench.py --batch $b --steps 8 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
}
run b7 7 M2D_X=0
run b7_pmin148 7 M2D_HALO_PERSIST_MIN=148
run b7_nopersist 7 M2D_HALO_PERSIST=0
run b7_grubg0 7 M2D_GRU_BG=0
run b7_grubg4 7 M2D_GRU_BG=4
run b64 64 M2D_X=0
run b64_pmin148 64 M2D_HALO_PERSIST_MIN=148
run b512 512 M2D_X=0
run b512_nopersist 512 M2D_HALO_PERSIST=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_b*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        fam=d.get("kernel_families",{}).get("rowconv",{})
        print("%-44s value %8.3f ms/step %8.2f e2e %8.3f launches %5s roof %.4f rowconv ms %.2f"%(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches_per_step"), d.get("roofline",{}).get("frac",0), fam.get("ms",0)))
    except Exception as e: print(f, "unreadable", e)
PY
