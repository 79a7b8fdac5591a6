cd $GRAFT_REPO_ROOT
TAG=${1:-pp}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vit_prep -s 10 -c 1 -o gpurun_out/${TAG}_vit_prep python bench.py --workload full --streams 256 --steps 14 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/${TAG}_vit_prep.ncu-rep 2>&1 | head -34
