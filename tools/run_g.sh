timeout 600 python -m pytest tests -m gpu -q -x -k "grid or ring or mma or shift or peer" 2>&1 | tail -2
timeout 300 python bench.py --workload grid --steps 10 --warmup 3 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('grid', d['ms_per_step'], d['stage_ms'], d['value'], d['best'])"
