# ncu --set full of ONE launch of one kernel (regex $KERNEL, skipping $SKIP matches) in the default bench workload
cd $GRAFT_REPO_ROOT
TAG=${1:-r2n}
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:${KERNEL}" -s ${SKIP:-8} -c ${COUNT:-1} -o gpurun_out/${TAG} python bench.py --steps 6 --warmup 6 --e2e-steps 0 --no-cpu-baseline --no-spot-check --no-ofdm-leg ${BENCH_ARGS} > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
