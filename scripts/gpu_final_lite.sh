#!/bin/bash
# Refresh of the round's single-GPU numbers without the long ncu --set full captures:  TAG=r2_final bash scripts/gpu_final_lite.sh
TAG=${TAG:-r2_final}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 200 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_cpu.json 2> /dev/null
python bench.py --impl reference-gpu --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_reference_gpu.json 2> /dev/null
python bench.py --workload kurt_kd --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_kurtkd.json 2> gpurun_out/${TAG}_bench_kurtkd.err
python bench.py --model resnet34 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_r34.json 2> gpurun_out/${TAG}_bench_r34.err
python bench.py --model resnet34 --batch 512 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu > gpurun_out/${TAG}_bench_r34_b512.json 2> gpurun_out/${TAG}_bench_r34_b512.err
python bench.py --model resnet20 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_r20.json 2> gpurun_out/${TAG}_bench_r20.err
python scripts/kernel_bench.py --impl tc --out gpurun_out/${TAG}_kernels.json > gpurun_out/${TAG}_kernels.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_bench.log 2>&1
ls gpurun_out | grep -c ${TAG}
