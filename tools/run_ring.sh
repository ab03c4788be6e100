grid() { timeout 300 python bench.py --workload grid --steps 5 --warmup 3 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('grid', d['ms_per_step'], d['stage_ms'], d['value'], d['best'])"; tail -3 gpurun_out/err.txt | cut -c1-300; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "grid 413 phase"; grid
echo "grid 413 plain"; TDR_GRID_PHASE_LOG2=0 grid
echo "grid 412 phase"; TDR_MMA_RING_CFG=412 grid
echo "grid 114 phase"; TDR_MMA_RING_CFG=114 grid
