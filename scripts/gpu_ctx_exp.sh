# experiment: the streams of one GPU split over 1 / 2 / 4 contexts (each with its own main + channel CUDA stream)
cd $GRAFT_REPO_ROOT
for n in 1 2 4; do
  (timeout 600 python bench.py --contexts $n --no-cpu-baseline --e2e-steps 0 --no-spot-check --no-c32-leg 2> gpurun_out/ctx${n}.err | tail -1) > gpurun_out/ctx${n}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ctx${n}.json"))
    print("contexts", ${n}, "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), {k: round(v/d["steps"],4) for k,v in d["kernel_ms"].items()}, "ofdm_only", round(d.get("ofdm_only",{}).get("ms_per_step",0),4), "fibs", d["counters"]["fibs_crc_ok"], d["counters"]["fibs_total"], "rs_fail", d["counters"]["superframes_rs_fail"])
except Exception as e:
    print("contexts", ${n}, "failed", e); print(open("gpurun_out/ctx${n}.err").read()[-1500:])
PY
done
