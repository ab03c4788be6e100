#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests/test_gpu_bench_configs.py tests/test_gpu_parity.py -x -q -m gpu -k "track or gated or uninit or step or update or repeat" 2>&1 | tail -4
echo "== i8 tracking passes"; python tools/track_1m.py 2>&1 | tail -7 | tee $O/r02_tracking_1m_i8.txt
echo "== fp16 ring tracking"; TDR_MMA_I8=0 python tools/track_1m.py 2>&1 | tail -7 | tee $O/r02_tracking_1m_ring.txt
python tools/class_api_bench.py > $O/r02_class_api_b.json 2> $O/r02_class_api_b.err; tail -c 1500 $O/r02_class_api_b.json
