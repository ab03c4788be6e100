#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_properties.py -x -q -m gpu -k "normal or weights or update or step or resample or small or track or nan or NaN or prefix or exact or pose or cfg2" 2>&1 | tail -2
python bench.py --workload tracking --steps 200 --warmup 10 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['p50_update_ms'],d['stage_ms'],d['e2e']['p50_ms'],d.get('verified'))"
