# the driver's two bench lines on the final code
cd $GRAFT_REPO_ROOT
(timeout 900 python bench.py 2> gpurun_out/r2_bench_default.err | tail -1) > gpurun_out/r2_bench_default.json
(timeout 900 python bench.py --impl reference 2> gpurun_out/r2_bench_reference.err | tail -1) > gpurun_out/r2_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_default.json"))
print("default", d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["roofline"]["frac"], d["roofline"]["issue"]["alu_pipe_frac"], d["roofline"]["traffic"], d["clocks"], d.get("spot_check", {}).get("identical"))
r = json.load(open("gpurun_out/r2_bench_reference.json"))
print("reference", r["value"], r["ms_per_step"])
PY
