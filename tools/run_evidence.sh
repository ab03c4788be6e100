set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r01_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench_global.json 2> gpurun_out/bench_global.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --workload tracking --steps 200 --warmup 10 > gpurun_out/r01_bench_tracking.json 2> gpurun_out/bench_tracking.err
timeout 600 python bench.py --workload grid --steps 10 --warmup 3 > gpurun_out/r01_bench_grid.json 2> gpurun_out/bench_grid.err
timeout 600 python bench.py --workload refine --steps 5 --warmup 2 > gpurun_out/r01_bench_refine.json 2> gpurun_out/bench_refine.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_global.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_mma -s 1 -c 1 -f -o gpurun_out/r01_score_list_tmem python bench.py --steps 1 --warmup 1 --no-cpu --particles 500000 > gpurun_out/ncu_full_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_mma -s 1 -c 1 -f -o gpurun_out/r01_score_ring_grid python bench.py --workload grid --steps 1 --warmup 1 > gpurun_out/ncu_full_ring.log 2>&1
cat gpurun_out/r01_smoke.txt gpurun_out/r01_pytest_gpu.txt
