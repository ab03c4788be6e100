#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_gpu_vs_reference.py -x -q -m gpu -k "dist or edt or EDT or geo or layer or map or refine or fixture" 2>&1 | tail -3
for impl in 0 2; do
TDR_EDT_IMPL=$impl ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_edt_launches_impl$impl.csv python tools/time_edt.py > $O/r02_edt_time_impl$impl.txt 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('$O/r02_edt_launches_impl$impl.csv', errors='ignore')) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[1:]:
    if 'edt' in r[ik] or 'seeds' in r[ik]: print($impl, r[ik][:50], r[iv])
PY
done
TDR_EDT_IMPL=0 python tools/time_edt.py
