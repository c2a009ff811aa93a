#!/bin/bash
# ncu --set full over one timed step for this repo's main kernels; only the raw CSV comes back (reports are large).
TAG=${TAG:-st}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --profile-from-start off \
   -k regex:"${K:-tc_conv2_kernel|tc_wgrad_kernel|bn_reduce_kernel|bn_apply_add_pack_kernel|bn_bwd_pack_kernel}" -c ${C:-140} \
   -o /tmp/${TAG}_prof -f python bench.py --steps 1 --profile-mode > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_prof.ncu-rep gpurun_out/${TAG}_raw.csv
