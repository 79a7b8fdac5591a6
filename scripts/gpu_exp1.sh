cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_channel_gpu.py -q -k "short_and_general" 2>&1 | tail -2)
run() {
  (timeout 600 python bench.py --no-cpu-baseline --e2e-steps 0 --no-spot-check --steps 40 2>/dev/null | tail -1) > gpurun_out/exp_$1.json
  python - <<PY
import json
d=json.load(open("gpurun_out/exp_$1.json"))
print("$1", "ms/step", round(d["ms_per_step"],4), "ofdm_only", round(d.get("ofdm_only",{}).get("ms_per_step",0),4))
PY
}
run base
DABGPU_OVERLAP_GROUPS=1 run split_groups
