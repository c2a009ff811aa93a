#!/bin/bash
# First GPU runs for the next round (each is one A/B of an existing knob on the bench; ~1 GPU-minute):
#   bash scripts/round2_first_runs.sh
# 1. multi-image halo patches for the 7x7 layers' forward / dgrad (default off, unmeasured)
K="fwd_tc or backward_tc or unit" NO_BENCH=1 TAG=r2a BDBNN_TC_HALO_SMALL=1 bash scripts/gpu_quick.sh
VAR=BDBNN_TC_HALO_SMALL VALS="0 1" bash scripts/gpu_ab.sh
# 2. 256-wide N tiles: 0 = off, 1 = dgrad, 2 = fp8 forward, 3 = both (default)
VAR=BDBNN_TC_BN256 VALS="0 3" bash scripts/gpu_ab.sh
# 3. BN statistics in the conv epilogues vs a separate pass
VAR=BDBNN_CONV_STATS VALS="0 1" bash scripts/gpu_ab.sh
# 4. own 1x1 shortcut path vs cuDNN
VAR=BDBNN_SHORTCUT_TC VALS="0 1" bash scripts/gpu_ab.sh
