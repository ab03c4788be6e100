timeout 600 python -m pytest tests -m gpu -q -x -k "refine" 2>&1 | tail -3
timeout 900 python bench.py --workload refine --steps 5 --warmup 2 > gpurun_out/r01_bench_refine.json 2> gpurun_out/bench_refine.err
tail -n 3 gpurun_out/bench_refine.err | cut -c1-400
python - <<'PY'
import json
for l in open('gpurun_out/r01_bench_refine.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline'])
PY
