#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_parity_gpu.py -q -k "gru or stock_gradient or conv_fwd" 2>&1 | tail -n 3
SEL='tf32x3 and (conv_fwd or merged or epilogue or full_length or linear or windowed or gru or batchnorm or elementwise or adam or persistent)'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_ops_gpu.py -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck.log | tail -n 3
SEL2='tf32x3 and (conv_fwd or epilogue or merged or persistent or gru)'
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_ops_gpu.py -q -k "$SEL2" -p no:cacheprovider > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck.log; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck.log | tail -n 3
for b in 7 64 512; do for g in tf32x3 tf32bf16; do
  timeout 600 python bench.py --gemm $g --batch $b --steps 6 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset > gpurun_out/bench_${g}_b$b.json 2> gpurun_out/bench_x.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_tf32*_b*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        fam=d.get("kernel_families",{}).get("rowconv",{})
        print("%-44s value %8.3f ms/step %8.2f roof %.4f rowconv ms %.2f"%(f, d["value"], d["ms_per_step"], d.get("roofline",{}).get("frac",0), fam.get("ms",0)))
    except Exception as e: print(f, "unreadable", e)
PY
