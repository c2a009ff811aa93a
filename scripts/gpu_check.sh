#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench line, kernel micro-bench, ncu launch list + full capture.
# Usage (from repo root, on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > gpurun_out/${TAG}_cpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python scripts/kernel_bench.py --impl ${IMPL:-xnor} --out gpurun_out/${TAG}_kernels.json > gpurun_out/${TAG}_kernels.log 2>&1
timeout 1200 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit: $?" >> gpurun_out/${TAG}_bench.err
if [ -z "$NO_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off --import-source on -k regex:"${NCU_K:-binconv_fwd_xnor|act_pack}" -c ${NCU_C:-4} \
   -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
tail -5 gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json | head -c 3000; tail -3 gpurun_out/${TAG}_bench.err
