#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BDBNN_TC_TS256=1 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x > gpurun_out/ts_tests.log 2>&1; tail -3 gpurun_out/ts_tests.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/ts_$name.json 2> gpurun_out/ts_$name.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/ts_$name.json") if l.startswith("{")][-1])
ks={k["kernel"]:k["ms_per_step"] for k in d["kernels"]}
print("$name", d["value"], d["ms_per_step"], {k:ks[k] for k in ("binconv_dgrad_tc","binconv_fwd_tc8","binconv_fwd_tc","shortcut_dgrad_tc")})
PY
}
run ts1 BDBNN_TC_TS256=1
run ts2 BDBNN_TC_TS256=2
run ts128_4 BDBNN_TC_TS128=4
run ts128_2 BDBNN_TC_TS128=2
