#!/bin/bash
# quick GPU iteration: selected tests (-k "$K"), then a short bench.  Usage: K="stem or ede" TAG=q1 bash scripts/gpu_quick.sh
TAG=${TAG:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "${K:-stem}" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
if [ -z "$NO_BENCH" ]; then
  timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, d.get("e2e"))
    for k in d.get("kernels",[])[:14]: print(k["kernel"], k["ms_per_step"], k["launches_per_step"])
except Exception as e: print("no bench json", e)
PY
fi
