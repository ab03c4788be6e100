#!/bin/bash
# round 2 evidence run (1 GPU): full GPU test suite, default bench line, launch list of the same command, ncu --set full
# of the two dominant kernels (cfg3: k_score_mma_i8 default config; cfg4: k_score_mma ring kernel)
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu -x > $O/r02_pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -2 $O/r02_pytest_gpu.txt
python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "ref rc $?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; echo "smoke rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_global.csv python bench.py --steps 2 --warmup 1 --no-sub --no-cpu --no-verify > $O/r02_launches_global.log 2>&1; echo "launch list rc $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_score_mma_i8 --launch-skip 1 -c 1 -o $O/r02_i8_default -f python bench.py --steps 1 --warmup 1 --no-sub --no-cpu --no-verify > $O/r02_ncu_i8_default.log 2>&1; echo "ncu i8 rc $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_score_mma --launch-skip 1 -c 1 -o $O/r02_ring_grid -f python bench.py --workload grid --steps 1 --warmup 1 --no-cpu --no-verify > $O/r02_ncu_ring_grid.log 2>&1; echo "ncu ring rc $?"
ls -la $O/*.ncu-rep | tail -4
