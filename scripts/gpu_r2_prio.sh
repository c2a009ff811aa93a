#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/prio_$name.json 2> gpurun_out/prio_$name.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/prio_$name.json") if l.startswith("{")][-1])
print("$name", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
}
run base X=1
run side_hi BDBNN_SIDE_PRIO=-1
run main_hi BDBNN_GRAPH_PRIO=-1
run bn3 BDBNN_BN_BWD_PER_SM=3
run bn3_side_hi BDBNN_BN_BWD_PER_SM=3 BDBNN_SIDE_PRIO=-1
run bn2_side_hi BDBNN_BN_BWD_PER_SM=2 BDBNN_SIDE_PRIO=-1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --workload kurt_kd > gpurun_out/prio_kurtkd.json 2> gpurun_out/prio_kurtkd.err || tail -3 gpurun_out/prio_kurtkd.err
BDBNN_TEACHER_SIDE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --workload kurt_kd > gpurun_out/prio_kurtkd_noside.json 2> gpurun_out/prio_kurtkd_noside.err
python - <<PY
import json
for n in ("kurtkd","kurtkd_noside"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/prio_{n}.json") if l.startswith("{")][-1]); print(n, d["value"], d["ms_per_step"])
    except Exception as e: print(n, "failed", e)
PY
timeout 300 python -m pytest tests/test_gpu_graph.py tests/test_gpu_ref_train.py -m gpu -q -x > gpurun_out/prio_tests.log 2>&1; tail -4 gpurun_out/prio_tests.log
