#!/bin/bash
# 2 GPUs: shard check across the 65536 tracked-set threshold; grid with the in-library exchange vs the NCCL all-reduce
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29533 tools/shard_check.py --workload global --particles 40000 --scans 3 > gpurun_out/r02_shard_check_global_b.txt 2>&1; echo "rc=$?"
grep ranks gpurun_out/r02_shard_check_global_b.txt
for c in fused fused-nccl; do
timeout 300 $R --master-port 29534 bench.py --gpus 2 --workload grid --steps 20 --warmup 3 --grid-collective $c > gpurun_out/r02_grid_2gpu_$c.json 2> gpurun_out/r02_grid_2gpu_$c.err; echo "rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_grid_2gpu_$c.json').read().strip().splitlines()[-1]);print('$c', d['ms_per_step'],d['stage_ms'],d['e2e']['ms_per_step'],d['best'],d.get('verified'))"
done
tail -3 gpurun_out/r02_grid_2gpu_fused.err
