# experiment helper: for every library build under scripts/_variants/, run the default bench (no CPU arm, no e2e) and print the kernel times
cd $GRAFT_REPO_ROOT
cp sdrplusplus-dab-radio-plugin_b200/csrc/libdabgpu.so /tmp/libdabgpu_base.so
for v in base $(ls scripts/_variants/*.so 2>/dev/null); do
  if [ "$v" != "base" ]; then cp $v sdrplusplus-dab-radio-plugin_b200/csrc/libdabgpu.so; fi
  name=$(basename $v .so)
  (timeout 600 python bench.py --no-cpu-baseline --e2e-steps 0 --no-spot-check ${BENCH_ARGS} 2>/dev/null | tail -1) > gpurun_out/var_${name}.json
  python - <<PY
import json
d=json.load(open("gpurun_out/var_${name}.json"))
print("${name}", "ms/step", round(d["ms_per_step"],4), {k: round(v/d["steps"],4) for k,v in d["kernel_ms"].items()}, "ofdm_only", round(d.get("ofdm_only",{}).get("ms_per_step",0),4))
PY
done
cp /tmp/libdabgpu_base.so sdrplusplus-dab-radio-plugin_b200/csrc/libdabgpu.so
