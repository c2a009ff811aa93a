#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -q -x > gpurun_out/side_tests.log 2>&1; tail -15 gpurun_out/side_tests.log
for v in 1 0; do
  BDBNN_WGRAD_SIDE=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-gpu --no-secondary > gpurun_out/side_bench_$v.json 2> gpurun_out/side_bench_$v.err
  tail -3 gpurun_out/side_bench_$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/side_bench_$v.json") if l.startswith("{")][-1])
print("side=$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("launch"))
PY
done
python -c "
import torch,time
x=torch.empty(1<<28,dtype=torch.float32,device='cuda')
for _ in range(3): x.zero_()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.zero_()
e1.record(); torch.cuda.synchronize()
print('memset 1 GiB write-only GB/s', x.numel()*4*10/e0.elapsed_time(e1)/1e6)
y=torch.empty_like(x)
e0.record()
for _ in range(10): s=x.sum()
e1.record(); torch.cuda.synchronize()
print('sum 1 GiB read-only GB/s', x.numel()*4*10/e0.elapsed_time(e1)/1e6)
"
