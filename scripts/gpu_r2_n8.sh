#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
run() { # name, bench args...
  name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus $NG "$@" > gpurun_out/n${NG}_$name.json 2> gpurun_out/n${NG}_$name.err
  echo "$name rc=$?"; python -c "
import json,sys
t=[l for l in open('gpurun_out/n${NG}_$name.json') if l.startswith('{')]
d=json.loads(t[-1]); print('$name', d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['config']['launch'][:40])"
}
run overlap2 --steps 20 --warmup 5 --no-secondary
BDBNN_DDP_BUCKETS=0 run single --steps 10 --warmup 3 --no-secondary --no-e2e
run r34_b512 --model resnet34 --batch 512 --steps 10 --warmup 3 --no-secondary
