#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BDBNN_TC_C64=2 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
   --kernel-name-base demangled -k regex:"tc_conv64_kernel<.int.1, .bool.1" --launch-count 1 -o /tmp/c64bst -f \
   python bench.py --steps 1 --profile-mode > gpurun_out/c64bst_ncu.log 2>&1
tail -3 gpurun_out/c64bst_ncu.log
ncu -i /tmp/c64bst.ncu-rep --page source --csv > gpurun_out/c64bst_source.csv 2>/dev/null
ncu -i /tmp/c64bst.ncu-rep --page raw --csv > gpurun_out/c64bst_raw.csv 2>/dev/null
wc -l gpurun_out/c64bst_source.csv gpurun_out/c64bst_raw.csv
