#!/bin/bash
# ncu --set full on the per-layer micro-bench: K=<kernel regex> C=<count> S=<skip> TAG=<name>
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${K:-tc_conv2_kernel}" -s ${S:-0} -c ${C:-3} \
   -o gpurun_out/${TAG:-kb}_prof -f python scripts/kernel_bench.py --impl tc > gpurun_out/${TAG:-kb}_ncu.log 2>&1
tail -3 gpurun_out/${TAG:-kb}_ncu.log
