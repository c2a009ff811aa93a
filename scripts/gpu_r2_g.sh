#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2g_tests.log 2>&1; echo "all rc=$?"; tail -8 gpurun_out/r2g_tests.log
for v in 1 0; do
  BDBNN_BWD_STATS=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/r2g_bench_bs$v.json 2> gpurun_out/r2g_bench_bs$v.err
  tail -c 150 gpurun_out/r2g_bench_bs$v.json; tail -2 gpurun_out/r2g_bench_bs$v.err
done
