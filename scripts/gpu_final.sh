#!/bin/bash
# Final single-GPU evidence run of a round:  TAG=r2_final bash scripts/gpu_final.sh   (outputs in gpurun_out/<TAG>_*)
TAG=${TAG:-r2_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_cpu.json 2> /dev/null
python bench.py --impl reference-gpu --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_reference_gpu.json 2> /dev/null
python bench.py --workload kurt_kd --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_kurtkd.json 2> gpurun_out/${TAG}_bench_kurtkd.err
python bench.py --model resnet34 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_r34.json 2> gpurun_out/${TAG}_bench_r34.err
python bench.py --model resnet34 --batch 512 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu > gpurun_out/${TAG}_bench_r34_b512.json 2> gpurun_out/${TAG}_bench_r34_b512.err
python bench.py --model resnet20 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_r20.json 2> gpurun_out/${TAG}_bench_r20.err
python scripts/kernel_bench.py --impl tc --out gpurun_out/${TAG}_kernels.json > gpurun_out/${TAG}_kernels.log 2>&1
# launch list of one step (eager launches, one step under ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_bench.log 2>&1
# ncu --set full of the main kernels over one step
timeout 1500 ncu --set full --clock-control none --profile-from-start off \
   -k regex:"tc_conv2_kernel|tc_conv64_kernel|tc_wgrad_kernel|bn_reduce_kernel|bn_apply_add_pack_kernel|bn_bwd_pack_kernel|weight_pack_multi|weight_wt_multi|bn_pool" -c 180 \
   -o /tmp/${TAG}_prof -f python bench.py --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
# source-level capture of ONE layer1 forward launch of the pixel-N kernel (launch 0 is the stem variant)
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"tc_conv64_kernel" \
   --launch-skip 1 --launch-count 1 -o /tmp/${TAG}_l1fwd -f python bench.py --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_l1fwd.log 2>&1
ncu -i /tmp/${TAG}_l1fwd.ncu-rep --page source --csv > gpurun_out/${TAG}_l1fwd_source.csv 2>/dev/null
ncu -i /tmp/${TAG}_l1fwd.ncu-rep --page raw --csv > gpurun_out/${TAG}_l1fwd_raw.csv 2>/dev/null
# loss kernels (config 3)
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"kurt|kd_" -c 40 -o /tmp/${TAG}_loss -f \
   python bench.py --workload kurt_kd --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_loss.log 2>&1
ncu -i /tmp/${TAG}_loss.ncu-rep --page raw --csv > gpurun_out/${TAG}_loss_raw.csv 2>/dev/null
ls -la gpurun_out | grep ${TAG} | wc -l
