grid() { timeout 300 python bench.py --workload grid --steps 5 --warmup 3 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('grid', d['ms_per_step'], d['stage_ms'], d['value'], d['best'])"; tail -3 gpurun_out/err.txt | cut -c1-300; }
glob() { timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('global', d['ms_per_step'], d['stage_ms'])"; tail -3 gpurun_out/err.txt | cut -c1-300; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "grid 413"; grid
echo "grid 412"; TDR_MMA_RING_CFG=412 grid
echo "global default"; glob
echo "global T1R4"; TDR_MMA_TILES=1 TDR_MMA_SPLIT=4 glob
echo "global ring"; TDR_MMA_KERNEL=2 glob
