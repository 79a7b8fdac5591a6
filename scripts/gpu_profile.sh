# ncu evidence: launch list of our kernels (full chain, 256 streams) + --set full captures of the top kernels
cd $GRAFT_REPO_ROOT
TAG=${1:-rX}
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 900 python bench.py 2>&1 | tail -1) > gpurun_out/${TAG}_bench_default.json 2>&1
(timeout 600 python bench.py --workload full --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full.json 2>&1
(timeout 600 python bench.py --workload full --streams 1024 --steps 24 --no-cpu-baseline --e2e-steps 8 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full_1024.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/${TAG}_launches_full.csv python bench.py --workload full --streams 256 --steps 24 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ofdm_demod -s 8 -c 1 -o gpurun_out/${TAG}_ofdm_demod python bench.py --streams 256 --steps 6 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-channel-leg > gpurun_out/${TAG}_ncu_ofdm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_lanes -s 14 -c 1 -o gpurun_out/${TAG}_viterbi_lanes python bench.py --workload full --streams 256 --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_vit.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vit_prep -s 14 -c 1 -o gpurun_out/${TAG}_vit_prep python bench.py --workload full --streams 256 --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_prep.log 2>&1
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_bench_default.json gpurun_out/${TAG}_bench_full.json gpurun_out/${TAG}_bench_full_1024.json
