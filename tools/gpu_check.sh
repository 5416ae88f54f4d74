#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench lines (default gemm mode + variants), reference arm,
# ncu launch list.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --gemm tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --batch 64 --steps 5 --warmup 3 --gemm tf32 --no-cpu-baseline > gpurun_out/bench_b64_tf32.json 2>> gpurun_out/bench.err
if [ "${1:-}" = "ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
fi
tail -n 3 gpurun_out/smoke.log; tail -n 12 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench.err
for f in bench bench_tf32 bench_b64 bench_b64_tf32; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "roof", d.get("roofline",{}).get("kernel"), round(d.get("roofline",{}).get("frac",0),4), "cpu", d.get("cpu_baseline",{}).get("value"))
    for k,v in sorted(d.get("kernel_families",{}).items(), key=lambda kv:-kv[1]["ms"]):
        print("   %-14s n=%5d ms=%8.3f tflops=%7.2f gbs=%8.1f share=%.3f"%(k,v["launches"],v["ms"],v["tflops"],v["gbs"],v["share_of_eager_step"]))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
