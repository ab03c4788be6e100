timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mma" 2>&1 | tail -2
for T in 4 2 1; do for ST in 9 10 12; do
  echo "T=$T ST=$ST"; TDR_MMA_TILES=$T TDR_MMA_ST_SHIFT=$ST timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['stage_ms'])"
done; done
