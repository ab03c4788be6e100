#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -x -q -m gpu -k "step or update or bench or gated or uninit or repeat or mma or track" 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 20 --warmup 3 --no-sub --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['verified'],d['gpu_launches'])"; done
