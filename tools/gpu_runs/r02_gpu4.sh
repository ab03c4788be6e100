#!/bin/bash
set -u
O=gpurun_out
echo "== pytest bench configs + mma"; timeout 1500 python -m pytest tests/test_gpu_bench_configs.py tests/test_gpu_parity.py -x -q -k "bench_configs or mma or search or ring or gate or smoke" > $O/r02_pytest_i8b.txt 2>&1; echo "rc $?"; tail -8 $O/r02_pytest_i8b.txt
python tools/sweep_score.py TDR_MMA_I8_CFG=222,223,224,232,233,142,144 2>&1 | tee $O/r02_sweep_i8_b.txt
python tools/sweep_score.py TDR_MMA_I8_CFG=232 TDR_MMA_SORT=0,1 2>&1 | tee -a $O/r02_sweep_i8_b.txt
