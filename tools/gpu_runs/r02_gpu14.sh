#!/bin/bash
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29534 bench.py --gpus 8 --workload grid --steps 20 --warmup 3 > gpurun_out/r02_grid_8gpu_fused_runs.json 2> gpurun_out/r02_grid_8gpu_fused_runs.err; echo "rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_grid_8gpu_fused_runs.json').read().strip().splitlines()[-1]);print('runs', d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['best'],d.get('verified'))"
tail -3 gpurun_out/r02_grid_8gpu_fused_runs.err
