#!/bin/bash
echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -8
for H in 1 0; do echo "=== wgrad bench HALO=$H"; BDBNN_WGRAD_HALO=$H timeout 600 python scripts/kernel_bench.py --impl tc --kernels wgrad_tc 2>&1 | grep -E "wgrad_tc" | cut -c1-150; done
