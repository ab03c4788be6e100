#!/bin/bash
set -u
O=gpurun_out
timeout 300 python tools/sweep_score.py TDR_MMA_I8_CFG=141,132,122,123,121,131 2>&1 | grep score | tee $O/r02_sweep_i8_k.txt
python bench.py --steps 20 --warmup 3 --no-sub --no-cpu > $O/r02_bench_e.json 2> $O/r02_bench_e.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_e.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['verified'])"
