set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r1_pytest.log 2>&1
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5) > gpurun_out/r1_smoke.log 2>&1
(timeout 600 python bench.py 2>&1 | tail -3) > gpurun_out/r1_bench_ofdm.log 2>&1
(timeout 600 python bench.py --workload full --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r1_bench_full.log 2>&1
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -3) > gpurun_out/r1_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_full.csv python bench.py --workload full --streams 256 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r1_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ofdm_demod -s 12 -c 2 -o gpurun_out/r1_ofdm_demod python bench.py --streams 256 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r1_ncu_ofdm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_viterbi -s 6 -c 2 -o gpurun_out/r1_viterbi python bench.py --workload full --streams 256 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r1_ncu_vit.log 2>&1
cat gpurun_out/r1_pytest.log gpurun_out/r1_smoke.log gpurun_out/r1_bench_ofdm.log gpurun_out/r1_bench_full.log gpurun_out/r1_bench_ref.log
