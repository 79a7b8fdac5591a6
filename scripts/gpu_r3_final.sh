# final pass of round 2 on the shipped tree: whole GPU suite, smoke(), memcheck of the ordered warp-kernel path, default bench line
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/r3f_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) >> gpurun_out/r3f_pytest.log
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest "tests/test_channel_gpu.py::test_viterbi_warp_kernel_longest_first_order" "tests/test_channel_gpu.py::test_viterbi_lane_plan_edge_cases" -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r3f_memcheck.log 2>&1
(timeout 900 python bench.py 2> gpurun_out/r3f_bench_default.err | tail -1) > gpurun_out/r3f_bench_default.json
cat gpurun_out/r3f_pytest.log gpurun_out/r3f_memcheck.log
python - <<PY
import json
d = json.load(open("gpurun_out/r3f_bench_default.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "spot", d["spot_check"]["identical"], "cpu", d["cpu_baseline"]["value"])
for k, v in d["channel_decode"].items():
    if isinstance(v, dict): print(k, round(v["ms_per_unit"], 4), round(v["viterbi_mbit_s"]))
PY
