#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python scripts/diag_stem.py > gpurun_out/diag_stem.log 2>&1; cat gpurun_out/diag_stem.log
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -q -x > gpurun_out/wg_tests.log 2>&1; tail -4 gpurun_out/wg_tests.log
for v in 0 2; do
  BDBNN_WG_STAGES=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/wg_bench_$v.json 2> gpurun_out/wg_bench_$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/wg_bench_$v.json") if l.startswith("{")][-1])
print("wg_stages=$v", d["value"], d["ms_per_step"], [ (k["kernel"], k["ms_per_step"]) for k in d["kernels"] if "wgrad" in k["kernel"]])
PY
done
