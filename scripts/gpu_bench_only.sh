# the default bench line on the current tree (all legs), plus the OFDM workload line
cd $GRAFT_REPO_ROOT
TAG=${1:-r3b}
(time timeout 900 python bench.py) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -4 gpurun_out/${TAG}_bench_default.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "spot", d["spot_check"]["identical"])
for k in ("ofdm_only", "ofdm_only_256", "ofdm_only_c32"):
    o = d.get(k) or {}
    print(k, o.get("ms_per_step"), o.get("iq_msps"), {a: round(b / d["steps"], 4) for a, b in (o.get("kernel_ms") or {}).items()}, (o.get("roofline") or {}).get("frac"))
PY
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_default.json").read().strip().splitlines()[-1])
print("channel_decode", json.dumps(d.get("channel_decode"), indent=1)[:2500])
PY
