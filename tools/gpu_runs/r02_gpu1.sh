#!/bin/bash
# round 2, GPU call 1: micro-benchmarks behind the score-kernel redesign, the new bench-config parity tests,
# compute-sanitizer on every mbarrier / TMEM pipeline, and the L1 request / sector / wavefront counters of the cfg3 kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/r02_gpu1_smi.txt 2>&1
echo "== gather_bench2"; timeout 300 tools/gather_bench2 > $O/r02_gather_bench2.txt 2>&1; echo "rc $?"
echo "== pytest bench configs"; timeout 900 python -m pytest tests/test_gpu_bench_configs.py -x -q > $O/r02_pytest_bench_configs.txt 2>&1; echo "rc $?"; tail -5 $O/r02_pytest_bench_configs.txt
echo "== sanitizer"
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $O/r02_sanitizer_$tool.log python tools/sanitize_case.py > $O/r02_sanitizer_$tool.stdout 2>&1
  echo "$tool exit $?"; tail -3 $O/r02_sanitizer_$tool.log
done
echo "== ncu counters: microbench patterns"
M=l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,smsp__inst_executed.sum,gpu__time_duration.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:k_ldg -c 120 --csv --log-file $O/r02_ncu_gather_bench2.csv tools/gather_bench2 > $O/r02_ncu_gather_bench2.log 2>&1; echo "rc $?"
echo "== ncu counters: cfg3 list kernel"
timeout 600 ncu --metrics $M --clock-control none -k regex:k_score_mma_list -c 2 --csv --log-file $O/r02_ncu_list_counters.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/r02_ncu_list_counters.log 2>&1; echo "rc $?"
echo "== bench (plain)"; timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02_bench_a.json 2> $O/r02_bench_a.err; echo "rc $?"; cat $O/r02_bench_a.json | head -c 1500
