#!/bin/bash
# State check of the whole tree on one B200, the way the driver runs it: smoke, all GPU tests, both bench arms.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log | cut -c1-300
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  .*Error|^FAILED|passed|failed" | tail -n 12
T0=$SECONDS
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$? wall $((SECONDS-T0)) s"
T0=$SECONDS
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "b200 arm rc=$? wall $((SECONDS-T0)) s"; tail -n 2 gpurun_out/bench.err | cut -c1-300
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1])
print("reference:", {k:r.get(k) for k in ("value","steps","warmup","ms_per_step","wall_s")}, r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"])
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("b200: value", round(d["value"],3), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "launches/step", d["gpu_launches_per_step"], "clocks", d["clocks"])
print("roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k not in ("measured_in","peak_source","traffic_source","tf32_peak_measured")})
print("tf32 peak", d["roofline"].get("tf32_peak_measured"))
print("cpu_baseline", d.get("cpu_baseline"))
print("torch_eager_gpu", json.dumps(d.get("torch_eager_gpu"))[:700])
t=d.get("throughput_regime") or {}
print("regime", {k:t.get(k) for k in ("batch_per_gpu","value","ms_per_step","sequences_per_s")}, {k:(round(v,4) if isinstance(v,float) else v) for k,v in (t.get("roofline") or {}).items() if k in ("kernel","achieved","peak","frac","traffic")})
print("e2e_device_dataset", {k:d["e2e_device_dataset"].get(k) for k in ("value","ms_per_step","h2d_bytes_per_step")})
for k,v in sorted(d.get("kernel_families",{}).items(), key=lambda kv:-kv[1]["ms"]):
    print("   %-14s n=%5d ms=%8.3f tflops=%7.2f gbs=%8.1f share=%.3f"%(k,v["launches"],v["ms"],v["tflops"],v["gbs"],v["share_of_eager_step"]))
PY
