#!/bin/bash
set -u
O=gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_bench_configs.py -x -q > $O/r02_pytest_i8d.txt 2>&1; echo "rc $?"; tail -5 $O/r02_pytest_i8d.txt
python tools/sweep_score.py TDR_MMA_TEX=0,1 TDR_MMA_I8_CFG=232,233,231,222 2>&1 | tee $O/r02_sweep_i8_d.txt
