#!/bin/bash
# stem on the pixel-N kernel: parity tests, then A/B bench (BDBNN_TC_C64_STEM=0/1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "stem" > gpurun_out/stem_tests.log 2>&1; tail -8 gpurun_out/stem_tests.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/stem_all_tests.log 2>&1; tail -5 gpurun_out/stem_all_tests.log
for v in 1 0; do
  BDBNN_TC_C64_STEM=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/stem_bench_$v.json 2> gpurun_out/stem_bench_$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/stem_bench_$v.json") if l.startswith("{")][-1])
print("stem=$v", d["value"], d["ms_per_step"], [ (k["kernel"], k["ms_per_step"]) for k in d.get("kernels", d.get("kernel_times", [])) if "stem" in k["kernel"]])
PY
done
