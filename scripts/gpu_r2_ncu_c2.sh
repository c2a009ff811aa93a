#!/bin/bash
# source-level captures of one fp8 forward and one dgrad(+BN sums) launch of the pixel-M conv kernel (layers 3/4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BDBNN_WGRAD_SIDE=0 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
   --kernel-name-base demangled -k regex:"tc_conv2_kernel<.int.0" --launch-skip 8 --launch-count 1 -o /tmp/c2f -f \
   python bench.py --steps 1 --profile-mode > gpurun_out/c2f_ncu.log 2>&1
ncu -i /tmp/c2f.ncu-rep --page source --csv > gpurun_out/c2f_source.csv 2>/dev/null
ncu -i /tmp/c2f.ncu-rep --page raw --csv > gpurun_out/c2f_raw.csv 2>/dev/null
BDBNN_WGRAD_SIDE=0 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
   --kernel-name-base demangled -k regex:"tc_conv2_kernel<.int.1, .int.1, .bool.1" --launch-skip 1 --launch-count 1 -o /tmp/c2d -f \
   python bench.py --steps 1 --profile-mode > gpurun_out/c2d_ncu.log 2>&1
ncu -i /tmp/c2d.ncu-rep --page source --csv > gpurun_out/c2d_source.csv 2>/dev/null
ncu -i /tmp/c2d.ncu-rep --page raw --csv > gpurun_out/c2d_raw.csv 2>/dev/null
wc -l gpurun_out/c2f_source.csv gpurun_out/c2d_source.csv
