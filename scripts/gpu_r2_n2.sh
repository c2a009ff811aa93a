#!/bin/bash
# 2-GPU run: NCCL correctness test + bench at N=2 (CUDA graph with the captured all-reduce, then eager)
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_nccl.py -q > gpurun_out/n2_nccl_tests.log 2>&1; echo "nccl tests rc=$?"; tail -6 gpurun_out/n2_nccl_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench_graph.json 2> gpurun_out/n2_bench_graph.err; echo "graph rc=$?"; tail -c 400 gpurun_out/n2_bench_graph.json; tail -5 gpurun_out/n2_bench_graph.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-graph --no-secondary > gpurun_out/n2_bench_eager.json 2> gpurun_out/n2_bench_eager.err; echo "eager rc=$?"; tail -c 300 gpurun_out/n2_bench_eager.json; tail -3 gpurun_out/n2_bench_eager.err
