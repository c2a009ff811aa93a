#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > gpurun_out/n2_final.json 2> gpurun_out/n2_final.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/n2_final.json") if l.startswith("{")][-1])
print("N=2", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
