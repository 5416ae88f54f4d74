#!/bin/bash
# round 2, call a: state of the round-1 tree with the new measurement legs + in-graph timelines + split modes at large batch
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 300 python tools/step_timeline.py 7 gpurun_out/timeline_b7.json > gpurun_out/timeline_b7.txt 2>&1; cat gpurun_out/timeline_b7.txt
timeout 300 python tools/step_timeline.py 64 gpurun_out/timeline_b64.json > gpurun_out/timeline_b64.txt 2>&1; tail -n 70 gpurun_out/timeline_b64.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b7.json 2> gpurun_out/bench_b7.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_b7.err
for b in 64 512; do
  for g in tf32x3 tf32bf16; do
    timeout 400 python bench.py --gemm $g --batch $b --steps 3 --warmup 3 --sub --no-cpu-baseline --no-device-dataset > gpurun_out/bench_${g}_b$b.json 2> gpurun_out/bench_mixed.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches", d.get("gpu_launches_per_step"), "roof", d.get("roofline",{}).get("kernel"), round(d.get("roofline",{}).get("frac",0),4), "peak", d.get("roofline",{}).get("peak"))
        for k in ("cpu_baseline","torch_eager_gpu"):
            if k in d: print("   ",k, json.dumps(d[k])[:600])
        if "throughput_regime" in d: print("    regime", {k:d["throughput_regime"].get(k) for k in ("batch_per_gpu","value","ms_per_step","sequences_per_s")}, (d["throughput_regime"].get("roofline") or {}).get("frac"))
        for k,v in sorted(d.get("kernel_families",{}).items(), key=lambda kv:-kv[1]["ms"]):
            print("   %-14s n=%5d ms=%8.3f tflops=%7.2f gbs=%8.1f share=%.3f"%(k,v["launches"],v["ms"],v["tflops"],v["gbs"],v["share_of_eager_step"]))
    except Exception as e: print(f, "unreadable", e)
PY
