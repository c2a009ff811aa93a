#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/r2d_tc.log 2>&1; echo "tc rc=$?"; tail -8 gpurun_out/r2d_tc.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2d_tests.log 2>&1; echo "all rc=$?"; tail -8 gpurun_out/r2d_tests.log
for cg in 1 0; do
  BDBNN_TC_CG2=$cg timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/r2d_bench_cg$cg.json 2> gpurun_out/r2d_bench_cg$cg.err
  tail -c 150 gpurun_out/r2d_bench_cg$cg.json; tail -2 gpurun_out/r2d_bench_cg$cg.err
done
BDBNN_TC_CG2=1 timeout 600 python scripts/kernel_bench.py --impl tc --out gpurun_out/r2d_kernels_cg1.json > gpurun_out/r2d_kernels_cg1.log 2>&1
BDBNN_TC_CG2=0 timeout 600 python scripts/kernel_bench.py --impl tc --out gpurun_out/r2d_kernels_cg0.json > gpurun_out/r2d_kernels_cg0.log 2>&1
