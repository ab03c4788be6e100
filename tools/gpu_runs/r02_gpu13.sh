#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -x -q -m gpu -k "grid or ring" 2>&1 | tail -3
for r in 1 8; do python tools/grid_shard_probe.py --ranks $r; done 2>&1 | grep ranks
python bench.py --workload grid --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['ms_per_step'],d['stage_ms'],d['best'],d.get('verified'))"
