#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BDBNN_TC_HALO_SMALL=1 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ref_train.py -m gpu -q -x > gpurun_out/hs_tests.log 2>&1; tail -3 gpurun_out/hs_tests.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/hs_$name.json 2> gpurun_out/hs_$name.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/hs_$name.json") if l.startswith("{")][-1])
ks={k["kernel"]:k["ms_per_step"] for k in d["kernels"]}
print("$name", d["value"], d["ms_per_step"], {k:ks[k] for k in ("binconv_dgrad_tc","binconv_fwd_tc8","binconv_fwd_tc","shortcut_dgrad_tc")})
PY
}
run hs1 BDBNN_TC_HALO_SMALL=1
run hs0 BDBNN_TC_HALO_SMALL=0
