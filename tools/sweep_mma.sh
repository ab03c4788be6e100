timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('global', d['ms_per_step'], d['stage_ms'])"
timeout 300 python bench.py --workload grid --steps 5 --warmup 3 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('grid', d['ms_per_step'], d['stage_ms'], d['value'], d['best'])"
TDR_MMA_KERNEL=1 timeout 300 python bench.py --workload grid --steps 3 --warmup 3 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('grid(list kernel)', d['ms_per_step'], d['stage_ms'], d['value'], d['best'])"
