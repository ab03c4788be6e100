#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_properties.py -x -q -m gpu -k "normal or weights or update or step or resample or bench or gated or nan or NaN or prefix or exact or pose" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-sub --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['verified'])"
python tools/norm_profile.py 2>/dev/null | tail -4
