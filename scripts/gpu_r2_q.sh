#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2r_tests.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/r2r_tests.log | cut -c1-300
for v in 1 2 0; do
  BDBNN_TC_C64=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/r2r_bench_c64_$v.json 2> gpurun_out/r2r_bench_c64_$v.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench_c64_$v.json') if l.startswith('{')][-1]); print('C64=$v', d['value'], d['ms_per_step'])
for k in d['kernels'][:7]: print('   ', k['kernel'], k['ms_per_step'])"
  tail -2 gpurun_out/r2r_bench_c64_$v.err
done
BDBNN_BWD_STATS=0 BDBNN_TC_C64=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary --no-e2e > gpurun_out/r2r_bench_nostats.json 2> /dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench_nostats.json') if l.startswith('{')][-1]); print('C64=1 BWD_STATS=0', d['value'], d['ms_per_step'])
for k in d['kernels'][:7]: print('   ', k['kernel'], k['ms_per_step'])"
