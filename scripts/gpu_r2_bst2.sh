#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/diag_c64bst.py > gpurun_out/diag_c64bst.log 2>&1; cat gpurun_out/diag_c64bst.log
for v in noh2d nod2d; do
  BDBNN_E2E_DIAG=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/e2e_$v.json 2> gpurun_out/e2e_$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/e2e_$v.json") if l.startswith("{")][-1])
print("e2e diag $v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done
