#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
for T in 2 4; do echo "=== TS128=$T"; BDBNN_TC_TS128=$T timeout 600 python scripts/kernel_bench.py --impl tc --kernels fwd_tc,dgrad_tc --layers layer2,layer3,layer4,layer2.0.conv1 2>&1 | grep -E "fwd_tc|dgrad_tc" | cut -c1-120; done
