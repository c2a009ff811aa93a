#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-secondary > gpurun_out/n2b_$name.json 2> gpurun_out/n2b_$name.err
  echo "$name rc=$?"; python -c "
import json,sys
t=[l for l in open('gpurun_out/n2b_$name.json') if l.startswith('{')]
d=json.loads(t[-1]); print('$name', d['value'], d['ms_per_step'], d['config']['launch'][:40])
for k in d['kernels'][:3]: print('   ', k['kernel'], k['ms_per_step'])"
}
run overlap4 BDBNN_DDP_BUCKETS=4
run overlap4_cta4 BDBNN_DDP_BUCKETS=4 NCCL_MAX_CTAS=4
run single BDBNN_DDP_BUCKETS=0
run single_cta8 BDBNN_DDP_BUCKETS=0 NCCL_MAX_CTAS=8
run overlap2 BDBNN_DDP_BUCKETS=2
