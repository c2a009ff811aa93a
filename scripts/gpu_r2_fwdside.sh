#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/fs_$name.json 2> gpurun_out/fs_$name.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/fs_$name.json") if l.startswith("{")][-1])
print("$name", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
}
run fwd1 X=1
run fwd0 BDBNN_FWD_SIDE=0
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/fs_tests.log 2>&1; tail -4 gpurun_out/fs_tests.log
