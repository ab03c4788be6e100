#!/bin/bash
# final single-GPU check of round 2: full GPU suite, smoke, default bench line
set -u
O=gpurun_out
python -m pytest tests -q -m gpu -x > $O/r02_pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -3 $O/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; echo "smoke rc $?"; tail -1 $O/r02_smoke.txt
python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc $?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['verified'],d['tracking']['p50_update_ms'],d['grid']['ms_per_step'])"
