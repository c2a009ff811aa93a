#!/bin/bash
# 2 GPUs: NCCL tests (side-stream wgrad folded into the all-reduce buckets), graph tests, N=2 bench side on/off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_graph.py -m gpu -q -x > gpurun_out/side2_tests.log 2>&1; tail -8 gpurun_out/side2_tests.log
for v in 1 0; do
  BDBNN_WGRAD_SIDE=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/side2_bench_$v.json 2> gpurun_out/side2_bench_$v.err
  tail -2 gpurun_out/side2_bench_$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/side2_bench_$v.json") if l.startswith("{")][-1])
print("N=2 side=$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("launch"))
PY
done
