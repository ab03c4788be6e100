#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -x -q -m gpu -k "mma or search or track or gated or uninit or i8 or counts" 2>&1 | tail -3
(timeout 300 python tools/sweep_score.py TDR_MMA_SORT=0,1 TDR_MMA_ST_SHIFT=8,10 2>&1 | grep score) | tee $O/r02_sweep_i8_o.txt
