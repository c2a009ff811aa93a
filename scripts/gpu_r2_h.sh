#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -k "bn_backward_sums or fused" > gpurun_out/r2h_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2h_tests.log
for v in 1 0; do
  BDBNN_BWD_STATS=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/r2h_bench_bs$v.json 2> gpurun_out/r2h_bench_bs$v.err
  tail -c 100 gpurun_out/r2h_bench_bs$v.json; tail -2 gpurun_out/r2h_bench_bs$v.err
done
