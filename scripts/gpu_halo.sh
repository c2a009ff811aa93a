#!/bin/bash
# halo-descriptor experiment: parity + layer timings for BDBNN_TC_HALO = 0 (off) / 1 (base_offset) / 2 (no base_offset)
mkdir -p gpurun_out
for H in 1 2 0; do
  echo "=== BDBNN_TC_HALO=$H"
  BDBNN_TC_HALO=$H timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4
done
for H in 1 0; do
  echo "=== bench HALO=$H"
  BDBNN_TC_HALO=$H timeout 600 python scripts/kernel_bench.py --impl tc 2>&1 | grep -E "fwd_tc|dgrad_tc" | cut -c1-150
done
