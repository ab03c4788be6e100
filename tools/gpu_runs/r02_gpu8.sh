#!/bin/bash
# round 2, call 8 (2 GPUs): library-level sharded filter — parity against one GPU, then the default bench at N = 2
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29533 tools/shard_check.py --workload global --particles 20000 > gpurun_out/r02_shard_check_global.txt 2>&1; echo "rc=$?"
tail -3 gpurun_out/r02_shard_check_global.txt
timeout 300 $R --master-port 29534 tools/shard_check.py --workload tracking --particles 20000 > gpurun_out/r02_shard_check_tracking.txt 2>&1; echo "rc=$?"
tail -3 gpurun_out/r02_shard_check_tracking.txt
timeout 600 $R --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 3 --no-sub > gpurun_out/r02_bench_2gpu_lib.json 2> gpurun_out/r02_bench_2gpu_lib.err; echo "rc=$?"
tail -c 3000 gpurun_out/r02_bench_2gpu_lib.json; tail -5 gpurun_out/r02_bench_2gpu_lib.err
timeout 600 $R --master-port 29536 bench.py --gpus 2 --steps 20 --warmup 3 --no-sub --no-verify --shard-impl torch > gpurun_out/r02_bench_2gpu_torch.json 2> gpurun_out/r02_bench_2gpu_torch.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("lib","torch"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_2gpu_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["stage_ms"], d["e2e"]["ms_per_step"], d.get("multi_gpu_check"))
    except Exception as e: print(f, "ERR", e)
PY
