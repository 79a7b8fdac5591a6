cd $GRAFT_REPO_ROOT
TAG=${1:-cp}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ofdm_ctl -s 12 -c 3 -o gpurun_out/${TAG}_ofdm_ctl python bench.py --streams 256 --steps 6 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-channel-leg > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/${TAG}_ofdm_ctl.ncu-rep 2>&1 | head -16
