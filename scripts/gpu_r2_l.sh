#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_c64.py > gpurun_out/r2l_diag.log 2>&1; tail -30 gpurun_out/r2l_diag.log
for dbg in 0 1 2 4 6; do
  echo "== C64 DBG=$dbg"
  BDBNN_TC_C64=1 BDBNN_TC_DBG=$dbg timeout 200 python scripts/kernel_bench.py --impl tc --layers layer1 --kernels fwd_tc,dgrad_tc 2>&1 | grep "fwd_tc\|dgrad_tc" | sed "s/'alg_MB.*TFLOPs'/TF/"
done
