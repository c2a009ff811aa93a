#!/bin/bash
mkdir -p gpurun_out
BDBNN_TC_C64=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"tc_conv64_kernel" --launch-skip 2 --launch-count 1 -o /tmp/c64prof -f python scripts/kernel_bench.py --impl tc --layers layer1 --kernels fwd_tc > gpurun_out/r2o_ncu.log 2>&1
ncu -i /tmp/c64prof.ncu-rep --page source --csv > gpurun_out/r2o_c64_source.csv 2>/dev/null
ncu -i /tmp/c64prof.ncu-rep --page raw --csv > gpurun_out/r2o_c64_raw.csv 2>/dev/null
ls -la gpurun_out/r2o_*; tail -3 gpurun_out/r2o_ncu.log
