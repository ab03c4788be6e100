#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -x -q -m gpu -k "mma or search or score or track or gated or uninit or step or update or repeat or counts or i8" 2>&1 | tail -4
timeout 300 python tools/sweep_score.py TDR_MMA_I8_CFG=141,131,151,231,222 2>&1 | grep score | tee $O/r02_sweep_i8_j.txt
