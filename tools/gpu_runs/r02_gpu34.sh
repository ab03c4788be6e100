#!/bin/bash
# 2-GPU box: the two tests the 1-GPU runs skip, plus a sanity pass over the final library build
set -u
O=gpurun_out
python -m pytest tests/test_gpu_bench_configs.py -q -m gpu -k "library_sharded or search_results_repeat or large_tracked" > $O/r02_pytest_gpu_2gpu.txt 2>&1; echo "pytest rc $?"; tail -3 $O/r02_pytest_gpu_2gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
