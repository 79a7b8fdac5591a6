# experiment: co-residency of k_viterbi_lanes (channel stream) and k_ofdm_demod2 (main stream)
cd $GRAFT_REPO_ROOT
run() {
  (timeout 600 python bench.py --no-cpu-baseline --e2e-steps 0 --no-spot-check --no-ofdm-leg --steps 60 2>/dev/null | tail -1) > gpurun_out/co_$1.json
  python - <<PY
import json
d=json.load(open("gpurun_out/co_$1.json"))
print("$1", "ms/step", round(d["ms_per_step"],4), {k: round(v/d["steps"],4) for k,v in d["kernel_ms"].items()})
PY
}
DABGPU_LANES_CTAS=2 DABGPU_DEMOD_CTAS=2 run l2_d2
DABGPU_LANES_CTAS=3 DABGPU_DEMOD_CTAS=1 run l3_d1
DABGPU_LANES_CTAS=3 DABGPU_DEMOD_CTAS=2 run l3_d2
DABGPU_LANES_CTAS=3 DABGPU_DEMOD_CTAS=3 run l3_d3
DABGPU_LANES_CTAS=3 DABGPU_DEMOD_CTAS=4 run l3_d4
DABGPU_LANES_CTAS=4 DABGPU_DEMOD_CTAS=4 run l4_d4
DABGPU_LANES_CTAS=2 DABGPU_DEMOD_CTAS=3 run l2_d3
DABGPU_CHAN_INLINE=1 run inline
