#!/bin/bash
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $R --master-port 29534 bench.py --gpus 8 --workload grid --steps 20 --warmup 3 > gpurun_out/r02_grid_8gpu_$name.json 2> gpurun_out/r02_grid_8gpu_$name.err; echo "rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r02_grid_8gpu_$name.json').read().strip().splitlines()[-1]);print('$name', d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['best'],d.get('verified'))"
}
run rot1_hint0 TDR_GRID_ROTATE=1 TDR_GRID_STORE_HINT=0
run rot1_hint1 TDR_GRID_ROTATE=1 TDR_GRID_STORE_HINT=1
run rot0_hint1 TDR_GRID_ROTATE=0 TDR_GRID_STORE_HINT=1
tail -3 gpurun_out/r02_grid_8gpu_rot1_hint0.err
