#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/r2k_tc.log 2>&1; echo "tc rc=$?"; tail -12 gpurun_out/r2k_tc.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2k_tests.log 2>&1; echo "all rc=$?"; tail -8 gpurun_out/r2k_tests.log
for v in 1 0; do
  echo "== C64=$v"
  BDBNN_TC_C64=$v timeout 300 python scripts/kernel_bench.py --impl tc --layers layer1 --kernels fwd_tc,dgrad_tc 2>&1 | grep "fwd_tc\|dgrad_tc" | sed "s/'alg_MB.*TFLOPs'/TF/"
  BDBNN_TC_C64=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/r2k_bench_c64_$v.json 2> gpurun_out/r2k_bench_c64_$v.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2k_bench_c64_$v.json') if l.startswith('{')][-1]); print('C64=$v', d['value'], d['ms_per_step'])
for k in d['kernels'][:6]: print('   ', k['kernel'], k['ms_per_step'])"
  tail -2 gpurun_out/r2k_bench_c64_$v.err
done
