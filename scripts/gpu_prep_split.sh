# experiment: k_vit_prep group split / window size
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_channel_gpu.py tests/test_protection_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -3)
run() {
  (timeout 600 python bench.py --no-cpu-baseline --e2e-steps 0 --no-spot-check --no-ofdm-leg 2>/dev/null | tail -1) > gpurun_out/ps_$1.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ps_$1.json"))
print("$1", "ms/step", round(d["ms_per_step"],4), {k: round(v/d["steps"],4) for k,v in d["kernel_ms"].items()})
PY
}
cp sdrplusplus-dab-radio-plugin_b200/csrc/libdabgpu.so /tmp/base.so
for sp in 1 2 4; do DABGPU_PREP_SPLIT=$sp run unit128_split$sp; done
run unit128_auto
cp scripts/_variants/unit64.so sdrplusplus-dab-radio-plugin_b200/csrc/libdabgpu.so
for sp in 1 2 4; do DABGPU_PREP_SPLIT=$sp run unit64_split$sp; done
(timeout 900 python -m pytest tests/test_channel_gpu.py tests/test_protection_gpu.py -x -q 2>&1 | tail -2)
cp /tmp/base.so sdrplusplus-dab-radio-plugin_b200/csrc/libdabgpu.so
