# round-2 profile pass: launch list of steady-state steps + ncu --set full of one launch of each hot kernel (default bench workload)
cd $GRAFT_REPO_ROOT
TAG=${1:-r2p}
bash scripts/gpu_r2_launches.sh ${TAG}
for K in k_viterbi_lanes k_ofdm_demod2 k_ofdm_ctl k_vit_prep k_chan_deinterleave k_dabplus; do
  KERNEL=$K SKIP=${SKIP:-8} bash scripts/gpu_r2_ncu1.sh ${TAG}_${K}
done
ls -la gpurun_out | tail -12
