run() { timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/err.txt | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('global', d['ms_per_step'], d['stage_ms'])"; tail -3 gpurun_out/err.txt | cut -c1-300; }
for cfg in "1 4" "1 2" "2 2" "2 1"; do set -- $cfg; echo "ATM T=$1 R=$2"; TDR_MMA_A_TMEM=1 TDR_MMA_TILES=$1 TDR_MMA_SPLIT=$2 run; 
TDR_MMA_A_TMEM=1 TDR_MMA_TILES=$1 TDR_MMA_SPLIT=$2 timeout 300 python -m pytest tests -m gpu -q -k "mma or search or global" 2>&1 | tail -8
done
