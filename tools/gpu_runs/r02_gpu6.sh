#!/bin/bash
set -u
O=gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_bench_configs.py tests/test_gpu_parity.py -x -q -k "bench_configs or mma or search or ring or gate or smoke or empty" > $O/r02_pytest_i8c.txt 2>&1; echo "rc $?"; tail -5 $O/r02_pytest_i8c.txt
python tools/sweep_score.py TDR_MMA_SKIP_RINGS=0,1 TDR_MMA_I8_CFG=232,233 2>&1 | tee $O/r02_sweep_i8_c.txt
