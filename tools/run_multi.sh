N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r01_bench_global_${N}gpu.json 2> gpurun_out/bench_global_g$N.err
timeout 600 $TR bench.py --gpus $N --workload grid --steps 10 --warmup 3 > gpurun_out/r01_bench_grid_${N}gpu_fused.json 2> gpurun_out/bench_grid_g$N.err
timeout 600 $TR bench.py --gpus $N --workload grid --grid-collective nccl --steps 10 --warmup 3 > gpurun_out/r01_bench_grid_${N}gpu_nccl.json 2> gpurun_out/bench_grid_nccl_g$N.err
for f in gpurun_out/r01_bench_global_${N}gpu.json gpurun_out/r01_bench_grid_${N}gpu_fused.json gpurun_out/r01_bench_grid_${N}gpu_nccl.json; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['n_gpus'], d['ms_per_step'], d['value'], d['stage_ms'], d['e2e'].get('ms_per_step'))
PY
done
tail -2 gpurun_out/bench_global_g$N.err gpurun_out/bench_grid_g$N.err
