#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1; echo "all rc=$?"; tail -6 gpurun_out/r2j_tests.log
for u in 2 4; do
  BDBNN_BN_UNROLL=$u timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/r2j_bench_u$u.json 2> gpurun_out/r2j_bench_u$u.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench_u$u.json') if l.startswith('{')][-1]); print('unroll $u', d['value'], d['ms_per_step'])
for k in d['kernels'][:5]: print('   ', k['kernel'], k['ms_per_step'])"
done
