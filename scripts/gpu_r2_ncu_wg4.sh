#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BDBNN_WGRAD_SIDE=0 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
   -k regex:"tc_wgrad_kernel" --launch-count 1 -o /tmp/wg4 -f \
   python bench.py --steps 1 --profile-mode > gpurun_out/wg4_ncu.log 2>&1
tail -2 gpurun_out/wg4_ncu.log
ncu -i /tmp/wg4.ncu-rep --page source --csv > gpurun_out/wg4_source.csv 2>/dev/null
ncu -i /tmp/wg4.ncu-rep --page raw --csv > gpurun_out/wg4_raw.csv 2>/dev/null
ncu -i /tmp/wg4.ncu-rep --page details --csv > gpurun_out/wg4_details.csv 2>/dev/null
wc -l gpurun_out/wg4_source.csv gpurun_out/wg4_raw.csv
