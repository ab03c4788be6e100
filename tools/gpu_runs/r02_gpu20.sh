#!/bin/bash
# 8 GPUs, final: sharded filter parity at 8 ranks (after the global-count kernel choice), then the default bench line
set -u
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29543 tools/shard_check.py --workload global --particles 20000 --scans 3 > $O/r02_shard_check_8gpu_global.txt 2>&1; echo "rc=$?"; grep ranks $O/r02_shard_check_8gpu_global.txt
timeout 300 $R --master-port 29544 tools/shard_check.py --workload tracking --particles 5000 --scans 3 > $O/r02_shard_check_8gpu_tracking.txt 2>&1; echo "rc=$?"; grep ranks $O/r02_shard_check_8gpu_tracking.txt
timeout 900 $R --master-port 29545 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_bench_8gpu_final.json 2> $O/r02_bench_8gpu_final.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_8gpu_final.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stage_ms"], "e2e", d["e2e"]["ms_per_step"], d.get("multi_gpu_check",{}).get("states_equal_single_gpu"))
print({k:v for k,v in d["grid"].items() if k not in ("verification","workload")})
PY
