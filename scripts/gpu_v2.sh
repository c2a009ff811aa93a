#!/bin/bash
mkdir -p gpurun_out
echo "=== V2 tests"
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -12
for V in 1 0; do
  echo "=== bench V2=$V"
  BDBNN_TC_V2=$V timeout 600 python scripts/kernel_bench.py --impl tc 2>&1 | grep -E "fwd_tc|dgrad_tc" | cut -c1-150
done
