#!/bin/bash
for D in 0 1 2 3 4 7; do
  echo "=== DBG=$D"
  BDBNN_TC_DBG=$D timeout 300 python scripts/kernel_bench.py --impl tc --layers layer1,layer2 --kernels fwd_tc,dgrad_tc 2>&1 | grep -E "fwd_tc|dgrad_tc" | cut -c1-120
done
