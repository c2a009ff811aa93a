#!/bin/bash
TAG=${1:-ll}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-400
