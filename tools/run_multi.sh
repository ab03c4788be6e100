N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r01_bench_global_${N}gpu.json 2> gpurun_out/bench_global_g$N.err
if [ "$2" != "noglobalonly" ]; then
timeout 600 $TR bench.py --gpus $N --workload grid --steps 10 --warmup 3 > gpurun_out/r01_bench_grid_${N}gpu_fused.json 2> gpurun_out/bench_grid_g$N.err
fi
for f in gpurun_out/r01_bench_global_${N}gpu.json gpurun_out/r01_bench_grid_${N}gpu_fused.json; do python - "$f" <<'PY'
import json,sys,os
if os.path.exists(sys.argv[1]):
  for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['n_gpus'], d['ms_per_step'], d['value'], d['stage_ms'], d['e2e'].get('ms_per_step'))
PY
done
