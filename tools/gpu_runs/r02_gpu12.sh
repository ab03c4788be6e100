#!/bin/bash
# 8 GPUs: cfg4 with the in-library exchange; the same with the NVLink cost stores switched off (what do they cost?)
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29534 bench.py --gpus 8 --workload grid --steps 20 --warmup 3 > gpurun_out/r02_grid_8gpu_fused.json 2> gpurun_out/r02_grid_8gpu_fused.err; echo "rc=$?"
TDR_GRID_SELF_ONLY=1 timeout 300 $R --master-port 29535 bench.py --gpus 8 --workload grid --steps 20 --warmup 3 --no-verify > gpurun_out/r02_grid_8gpu_selfonly.json 2> gpurun_out/r02_grid_8gpu_selfonly.err; echo "rc=$?"
timeout 300 $R --master-port 29536 bench.py --gpus 8 --workload grid --steps 20 --warmup 3 --grid-collective fused-nccl --no-verify > gpurun_out/r02_grid_8gpu_fusednccl.json 2> gpurun_out/r02_grid_8gpu_fusednccl.err; echo "rc=$?"
for c in fused selfonly fusednccl; do
python -c "
import json;d=json.loads(open('gpurun_out/r02_grid_8gpu_$c.json').read().strip().splitlines()[-1]);print('$c', d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['best'],d.get('verified'))"
done
tail -3 gpurun_out/r02_grid_8gpu_fused.err
