# round 2, first GPU pass: parity tests, the default bench (full chain, 1024 streams), launch list + one --set full capture of a whole step
cd $GRAFT_REPO_ROOT
TAG=${1:-r2a}
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 900 python bench.py 2> gpurun_out/${TAG}_bench_default.err | tail -1) > gpurun_out/${TAG}_bench_default.json
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1) > gpurun_out/${TAG}_bench_reference.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 130 -c 390 --csv --log-file gpurun_out/${TAG}_launches_full.csv python bench.py --steps 12 --warmup 6 --e2e-steps 0 --no-cpu-baseline --no-spot-check --no-ofdm-leg > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:k_ofdm_ctl|k_ofdm_demod2|k_vit_prep|k_viterbi_lanes|k_dabplus' -s 80 -c 8 -o gpurun_out/${TAG}_step python bench.py --steps 8 --warmup 6 --e2e-steps 0 --no-cpu-baseline --no-spot-check --no-ofdm-leg > gpurun_out/${TAG}_ncu_step.log 2>&1
tail -3 gpurun_out/${TAG}_bench_default.err
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_bench_default.json gpurun_out/${TAG}_bench_reference.json
