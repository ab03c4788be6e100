#!/bin/bash
# 4 GPUs: the one world size of the driver's scaling run that had not been exercised this round
set -u
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29545 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r02_bench_4gpu.json 2> $O/r02_bench_4gpu.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_4gpu.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["stage_ms"], "e2e", d["e2e"]["ms_per_step"], d.get("multi_gpu_check",{}).get("states_equal_single_gpu"), d["verified"])
print({k:v for k,v in d["grid"].items() if k in ("ms_per_step","vs_1gpu","verified","score_kernel_ms")})
PY
tail -3 $O/r02_bench_4gpu.err
