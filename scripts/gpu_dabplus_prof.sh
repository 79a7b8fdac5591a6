cd $GRAFT_REPO_ROOT
TAG=${1:-dpp}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dabplus -s 12 -c 1 -o gpurun_out/${TAG}_dabplus python bench.py --workload full --streams 256 --steps 16 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/${TAG}_dabplus.ncu-rep 2>&1 | head -42
