#!/bin/bash
mkdir -p gpurun_out
for pw in 0 1; do
  echo "== PW8=$pw"
  BDBNN_TC_PW8=$pw timeout 300 python scripts/kernel_bench.py --impl tc --layers layer1,layer2,layer3 --kernels fwd_tc,dgrad_tc,fwd_tc8 2>&1 | grep "fwd_tc\|dgrad_tc" | sed "s/'alg_MB.*TFLOPs'/TF/"
done > gpurun_out/r2f_pw8.log 2>&1
cat gpurun_out/r2f_pw8.log
BDBNN_TC_PW8=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/r2f_tc_pw8.log 2>&1; tail -3 gpurun_out/r2f_tc_pw8.log
BDBNN_TC_CG2=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "fwd_tc or backward_tc_vs" > gpurun_out/r2f_tc_cg2.log 2>&1; tail -3 gpurun_out/r2f_tc_cg2.log
