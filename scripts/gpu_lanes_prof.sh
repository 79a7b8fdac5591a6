# ncu evidence for the lane-per-trellis Viterbi path: launch list + --set full of k_viterbi_lanes and k_vit_prep
cd $GRAFT_REPO_ROOT
TAG=${1:-lp}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/${TAG}_launches_full.csv python bench.py --workload full --streams 256 --steps 16 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_lanes -s 14 -c 1 -o gpurun_out/${TAG}_viterbi_lanes python bench.py --workload full --streams 256 --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_lanes.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vit_prep -s 14 -c 1 -o gpurun_out/${TAG}_vit_prep python bench.py --workload full --streams 256 --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_prep.log 2>&1
python scripts/ncu_summary.py gpurun_out/${TAG}_viterbi_lanes.ncu-rep 2>&1 | head -60
tail -3 gpurun_out/${TAG}_ncu_lanes.log
