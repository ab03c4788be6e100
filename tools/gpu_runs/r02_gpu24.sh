#!/bin/bash
set -u
O=gpurun_out
timeout 300 python tools/sweep_score.py TDR_I8_DIAG_HALF_B=0,1 2>&1 | grep score | tee $O/r02_sweep_i8_l.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_score_mma_i8 --launch-skip 1 -c 1 -o $O/r02_i8_split -f python bench.py --steps 1 --warmup 1 --no-sub --no-cpu --no-verify > $O/r02_ncu_i8_split.log 2>&1; echo "ncu rc $?"
