#!/bin/bash
# which component bounds the CTA-pair kernel? BDBNN_TC_DBG: 1 = no global stores, 2 = no TMEM loads (no epilogue work), 4 = no MMAs
mkdir -p gpurun_out
for cg in 0 1; do for dbg in 0 1 2 4 6; do
  echo "== CG2=$cg DBG=$dbg"
  BDBNN_TC_CG2=$cg BDBNN_TC_DBG=$dbg timeout 300 python scripts/kernel_bench.py --impl tc --layers layer1,layer3 --kernels fwd_tc,dgrad_tc,fwd_tc8 2>&1 | grep "fwd_tc\|dgrad_tc" | sed "s/'alg_MB.*TFLOPs'/TF/"
done; done > gpurun_out/r2e_dbg.log 2>&1
tail -60 gpurun_out/r2e_dbg.log
