# longest-first order of the warp-per-trellis kernel: the Viterbi / channel / protection parity tests, then the channel-decode legs of the bench
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_channel_gpu.py tests/test_protection_gpu.py tests/test_golden_gpu.py tests/test_adapters_gpu.py -x -q 2>&1 | tail -6) > gpurun_out/ord_pytest.log
cat gpurun_out/ord_pytest.log
(timeout 900 python bench.py --no-cpu-baseline --e2e-steps 0 --no-spot-check --no-c32-leg --no-ofdm-leg 2> gpurun_out/ord_bench.err | tail -1) > gpurun_out/ord_bench.json
tail -n 3 gpurun_out/ord_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/ord_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"])
for k, v in d["channel_decode"].items():
    if isinstance(v, dict): print(k, round(v["ms_per_unit"], 4), round(v["viterbi_mbit_s"]), v["fibs_crc_ok"], v["fibs_total"])
PY
