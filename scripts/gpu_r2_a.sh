#!/bin/bash
# Round-2 first GPU run: full GPU test suite, default bench, kurt_kd bench (config 3) + ncu of the loss kernels,
# per-layer kernel bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
python bench.py --workload kurt_kd --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_kurtkd.json 2> gpurun_out/r2a_bench_kurtkd.err
python scripts/kernel_bench.py --impl tc --out gpurun_out/r2a_kernels.json > gpurun_out/r2a_kernels.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off \
   -k regex:"kurt|kd_" -c 40 -o /tmp/r2a_loss -f python bench.py --workload kurt_kd --steps 1 --profile-mode > gpurun_out/r2a_ncu_loss.log 2>&1
ncu -i /tmp/r2a_loss.ncu-rep --page raw --csv > gpurun_out/r2a_loss_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
