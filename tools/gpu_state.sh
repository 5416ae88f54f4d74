#!/bin/bash
# State check of the whole tree on one B200: smoke, all GPU tests, bench (both arms), batch-64 bench, launch list.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  .*Error|^FAILED|passed|failed" > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64.json 2>> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 4200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
tail -n 3 gpurun_out/smoke.log; tail -n 25 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json
for f in bench bench_b64; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches/step", d["gpu_launches_per_step"], "roof", d.get("roofline",{}).get("kernel"), round(d.get("roofline",{}).get("frac",0),4), "cpu", d.get("cpu_baseline",{}).get("value"))
    for k,v in sorted(d.get("kernel_families",{}).items(), key=lambda kv:-kv[1]["ms"]):
        print("   %-14s n=%5d ms=%8.3f tflops=%7.2f gbs=%8.1f share=%.3f"%(k,v["launches"],v["ms"],v["tflops"],v["gbs"],v["share_of_eager_step"]))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
