# round-2 closing check: the whole GPU suite, smoke(), the default bench line (all legs) and the reference arm
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/r3_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/r3_smoke.log
(time timeout 900 python bench.py) > gpurun_out/r3_bench_default.json 2> gpurun_out/r3_bench_default.err
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/r3_bench_reference.json 2> gpurun_out/r3_bench_reference.err
cat gpurun_out/r3_pytest.log gpurun_out/r3_smoke.log
tail -n 4 gpurun_out/r3_bench_default.err; tail -n 4 gpurun_out/r3_bench_reference.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3_bench_default.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("ofdm_only", d["ofdm_only"]["ms_per_step"], d["ofdm_only"]["roofline"]["frac"], d["ofdm_only"]["roofline"]["issue"])
print("c32", json.dumps(d.get("ofdm_only_c32"))[:1500])
print("cpu", json.dumps(d.get("cpu_baseline"))[:1200])
PY
