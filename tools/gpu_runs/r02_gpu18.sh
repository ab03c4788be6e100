#!/bin/bash
set -u
O=gpurun_out
for band in 128 64 32 256; do
TDR_EDT_BAND=$band ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $O/r02_edt_launches_band$band.csv python tools/time_edt.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('$O/r02_edt_launches_band$band.csv', errors='ignore')) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[-6:]:
    if 'edt' in r[ik]: print($band, r[ik][:50], r[iv])
PY
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -x -q -m gpu -k "dist or edt or EDT or geo or layer" 2>&1 | tail -2
