# tuning sweep of the theta-search kernel on cfg3 (env knobs of tdr_create): super-tile side, column-segment width
run() { timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%.3f'%d['stage_ms']['score'])"; }
for st in 9 10 11 12; do for sg in 1 2 3; do echo -n "ST_SHIFT=$st SEG_SHIFT=$sg score_ms="; TDR_MMA_ST_SHIFT=$st TDR_MMA_SEG_SHIFT=$sg run; done; done
