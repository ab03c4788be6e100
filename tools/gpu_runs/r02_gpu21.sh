#!/bin/bash
set -u
O=gpurun_out
python tools/sweep_score.py TDR_MMA_I8_CFG=141,151 TDR_I8_DIAG_HALF_B=0,1 2>&1 | grep score | tee $O/r02_sweep_i8_i.txt
