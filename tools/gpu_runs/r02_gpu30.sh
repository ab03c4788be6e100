#!/bin/bash
set -u
O=gpurun_out
timeout 600 python bench.py --workload refine --steps 5 --warmup 3 > $O/r02_bench_refine.json 2> $O/r02_bench_refine.err; echo "rc $?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_refine.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d.get('stage_ms'),d['e2e'])"
timeout 300 python bench.py --workload tracking --steps 200 --warmup 10 > $O/r02_bench_tracking.json 2> $O/r02_bench_tracking.err; echo "rc $?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_tracking.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['p50_update_ms'],d['stage_ms'],d['e2e']['p50_ms'],d.get('verified'))"
