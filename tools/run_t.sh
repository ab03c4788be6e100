timeout 900 python -m pytest tests -m gpu -q -x -k "normal or resam or weight or small or update or step or host or track" 2>&1 | tail -3
timeout 300 python bench.py --workload tracking --steps 200 --warmup 10 --no-cpu 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('tracking', d['ms_per_step'], d['p50_update_ms'], d['stage_ms'], d['gpu_launches'], d['e2e']['p50_ms'])"
