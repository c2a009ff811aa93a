#!/bin/bash
# Focused GPU call for the tcgen05 kernels: parity tests + per-layer micro-bench.
TAG=${1:-tc}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py::test_kd_layer_matches_reference_golden -x -q 2>&1 | tail -60 > gpurun_out/${TAG}_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/kernel_bench.py --impl tc --out gpurun_out/${TAG}_kernels.json > gpurun_out/${TAG}_kernels.log 2>&1
tail -60 gpurun_out/${TAG}_pytest.log; grep -v "'wgrad'" gpurun_out/${TAG}_kernels.log | cut -c1-160
