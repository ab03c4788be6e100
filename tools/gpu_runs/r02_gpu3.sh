#!/bin/bash
set -u
O=gpurun_out
for cfg in 222 232; do
TDR_MMA_I8_CFG=$cfg timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_score_mma_i8 --launch-skip 2 -c 1 -o $O/r02_i8_$cfg -f python tools/sweep_score.py TDR_MMA_I8_CFG=$cfg --particles=500000 > $O/r02_ncu_i8_$cfg.log 2>&1; echo "rc $?"
done
ls -la $O/*.ncu-rep | tail -3
