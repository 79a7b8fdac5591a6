# quick iteration pass: GPU parity tests + the default bench line (no CPU arm)
cd $GRAFT_REPO_ROOT
TAG=${1:-r2q}
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 900 python bench.py --no-cpu-baseline ${BENCH_ARGS} 2> gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_pytest.log
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"])
print("kernel_ms", d["kernel_ms"])
print("spot", d.get("spot_check"))
print("ofdm_only", d.get("ofdm_only", {}).get("ms_per_step"), d.get("ofdm_only", {}).get("kernel_ms"))
print("counters", d.get("counters"))
PY
