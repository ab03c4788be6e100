#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest bench configs"; timeout 1200 python -m pytest tests/test_gpu_bench_configs.py -x -q > $O/r02_pytest_bench_configs2.txt 2>&1; echo "rc $?"; tail -15 $O/r02_pytest_bench_configs2.txt
echo "== pytest mma"; timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "mma or smoke or search or ring" > $O/r02_pytest_mma.txt 2>&1; echo "rc $?"; tail -8 $O/r02_pytest_mma.txt
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== bench TDR_MMA_I8=$1 TDR_MMA_SORT=$2"
  TDR_MMA_I8=$1 TDR_MMA_SORT=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02_bench_i8_$1_$2.json 2> $O/r02_bench_i8_$1_$2.err; echo "rc $?"
  python -c "import json;d=json.load(open('$O/r02_bench_i8_$1_$2.json'));print(d['ms_per_step'],d['stage_ms'])"
done
