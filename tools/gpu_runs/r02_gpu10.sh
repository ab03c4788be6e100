#!/bin/bash
set -x
mkdir -p gpurun_out
for r in 1 2 4 8 16; do python tools/grid_shard_probe.py --ranks $r; done > gpurun_out/r02_grid_shard_probe.txt 2>&1
cat gpurun_out/r02_grid_shard_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_grid_shard8_launches.csv python tools/grid_shard_probe.py --ranks 8 --steps 2 > /dev/null 2>&1
# normalise with the speculative sum chains: parity + timing
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -x -q -m gpu -k "normal or weights or update or step or resample or bench or gated" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-sub --no-cpu > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_d.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['verified'])"
