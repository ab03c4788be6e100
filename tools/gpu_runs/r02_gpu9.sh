#!/bin/bash
# round 2, call 9 (8 GPUs): library-level sharded filter at 8 ranks (parity), then the default bench line at N = 8
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29543 tools/shard_check.py --workload global --particles 20000 --scans 3 > gpurun_out/r02_shard_check_8.txt 2>&1; echo "rc=$?"
tail -2 gpurun_out/r02_shard_check_8.txt
timeout 900 $R --master-port 29545 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "rc=$?"
tail -5 gpurun_out/r02_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_8gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stage_ms"], "e2e", d["e2e"]["ms_per_step"], d.get("multi_gpu_check"))
print({k:v for k,v in d["grid"].items() if k not in ("verification","workload")})
PY
