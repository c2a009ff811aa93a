#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c2_tests.log 2>&1; tail -3 gpurun_out/c2_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/c2_bench.json") if l.startswith("{")][-1])
print("c2", d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], k["ms_per_step"]) for k in d["kernels"][:8]])
PY
