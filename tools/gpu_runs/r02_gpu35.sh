#!/bin/bash
set -u
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_global.csv python bench.py --steps 2 --warmup 1 --no-sub --no-cpu --no-verify > $O/r02_launches_global.log 2>&1; echo "launch list rc $?"
