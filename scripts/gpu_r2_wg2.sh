#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/wg2_tests.log 2>&1; tail -3 gpurun_out/wg2_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/wg2_bench.json 2> gpurun_out/wg2_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/wg2_bench.json") if l.startswith("{")][-1])
print("wg2", d["value"], d["ms_per_step"], d["e2e"]["value"], [(k["kernel"], k["ms_per_step"]) for k in d["kernels"] if "wgrad" in k["kernel"]])
PY
timeout 300 python scripts/kernel_bench.py --impl tc --out gpurun_out/wg2_kernels.json > gpurun_out/wg2_kernels.log 2>&1
python - <<PY
import json
for r in json.load(open("gpurun_out/wg2_kernels.json")):
    if "wgrad" in r["kernel"]: print(r["layer"], r["kernel"], r["ms"], r["TFLOPs"])
PY
