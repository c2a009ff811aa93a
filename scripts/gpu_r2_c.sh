#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; tail -12 gpurun_out/r2c_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 200 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
BDBNN_Y_I16=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/r2c_bench_noi16.json 2> gpurun_out/r2c_bench_noi16.err
python scripts/kernel_bench.py --impl tc --out gpurun_out/r2c_kernels.json > gpurun_out/r2c_kernels.log 2>&1
