timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mma or sharded" 2>&1 | tail -2
for cfg in "2 1 2" "1 4 2" "1 2 2" "4 1 2"; do set -- $cfg
  echo "T=$1 R=$2 SEG=$3"; TDR_MMA_TILES=$1 TDR_MMA_SPLIT=$2 TDR_MMA_SEG_SHIFT=$3 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['stage_ms']['score'])"
done
