#!/bin/bash
set -u
O=gpurun_out
(timeout 300 python tools/sweep_score.py TDR_MMA_SORT=1 TDR_MMA_ST_SHIFT=11,12,13,16 2>&1 | grep score
timeout 300 python tools/sweep_score.py TDR_MMA_SORT=0 TDR_MMA_ST_SHIFT=7,8 TDR_MMA_SEG_SHIFT=0,1,2 2>&1 | grep score) | tee $O/r02_sweep_i8_n.txt
