#!/bin/bash
# A/B of one environment switch at batch 7: tools/gpu_ab.sh VAR valA valB [batch]
set -u
mkdir -p gpurun_out
V=$1; A=$2; B=$3; BATCH=${4:-7}
EXTRA=${M2D_AB_EXTRA:-}
for rep in 1 2; do for val in $A $B; do
  env $EXTRA $V=$val timeout 400 python bench.py --batch $BATCH --steps 15 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-throughput-regime --no-device-dataset --no-roofline > gpurun_out/ab_${val}_$rep.json 2> gpurun_out/ab.err
  python - $V $val $rep <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/ab_{sys.argv[2]}_{sys.argv[3]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]}={sys.argv[2]} rep {sys.argv[3]}: value {d['value']:.3f} ms/step {d['ms_per_step']:.3f} launches {d['gpu_launches_per_step']}")
except Exception as e: print("unreadable", e, open("gpurun_out/ab.err").read()[-500:])
PY
done; done
