#!/bin/bash
# A/B of one environment knob on the bench: VAR=BDBNN_TC_BN256 bash scripts/gpu_ab.sh
mkdir -p gpurun_out
for v in ${VALS:-0 1}; do
  env ${VAR}=$v timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("${VAR}=$v", d["value"], d["ms_per_step"], {k["kernel"]: k["ms_per_step"] for k in d["kernels"][:8]})
PY
done
