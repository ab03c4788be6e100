#!/bin/bash
set -u
O=gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_score_mma_i8 --launch-skip 2 -c 1 -o $O/r02_i8_lean_232 -f python tools/sweep_score.py TDR_MMA_I8_CFG=232 --particles=500000 > $O/r02_ncu_i8_lean.log 2>&1; echo "rc $?"
