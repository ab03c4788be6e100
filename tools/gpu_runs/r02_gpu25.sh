#!/bin/bash
set -u
O=gpurun_out
(timeout 300 python tools/sweep_score.py TDR_MMA_SORT=0,1 TDR_MMA_ST_SHIFT=8,9,10 2>&1 | grep score
timeout 300 python tools/sweep_score.py TDR_MMA_SORT=0 TDR_MMA_SEG_SHIFT=2,3,4,5 2>&1 | grep score) | tee $O/r02_sweep_i8_m.txt
