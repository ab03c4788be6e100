timeout 900 python -m pytest tests -m gpu -q -x -k "normal or resam or weight or prefix or sharded or update or exact or pose or step or host" 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('global', d['ms_per_step'], d['stage_ms'])"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --particles 8000000 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('global 8M', d['ms_per_step'], d['stage_ms'])"
