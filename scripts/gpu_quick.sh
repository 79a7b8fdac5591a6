# quick GPU check: parity tests + both bench workloads (no ncu)
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -22) > gpurun_out/q_pytest.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/q_bench_ofdm.log 2>&1
(timeout 600 python bench.py --workload full --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/q_bench_full.log 2>&1
cat gpurun_out/q_pytest.log gpurun_out/q_bench_ofdm.log gpurun_out/q_bench_full.log
