#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bdbnn_b200/csrc -I include scripts/mma_probe.cu -o /tmp/mma_probe 2> /dev/null
timeout 120 /tmp/mma_probe > gpurun_out/mma_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/mma_probe.txt | head -80
