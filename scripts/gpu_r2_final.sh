# round-2 evidence pass: GPU tests, the driver's two bench lines, launch list, ncu --set full of one launch of each hot kernel
cd $GRAFT_REPO_ROOT
TAG=${1:-r2}
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 900 python bench.py 2> gpurun_out/${TAG}_bench_default.err | tail -1) > gpurun_out/${TAG}_bench_default.json
(timeout 900 python bench.py --impl reference 2> gpurun_out/${TAG}_bench_reference.err | tail -1) > gpurun_out/${TAG}_bench_reference.json
bash scripts/gpu_r2_launches.sh ${TAG} > gpurun_out/${TAG}_launches_summary.txt 2>&1
for K in k_viterbi_lanes k_ofdm_demod2 k_vit_prep k_dabplus; do
  KERNEL=$K SKIP=8 bash scripts/gpu_r2_ncu1.sh ${TAG}_${K} > /dev/null 2>&1
done
KERNEL=k_ofdm_ctl SKIP=18 COUNT=2 bash scripts/gpu_r2_ncu1.sh ${TAG}_k_ofdm_ctl > /dev/null 2>&1
cat gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_launches_summary.txt | head -14
python - <<PY
import json
for f in ("default", "reference"):
    d = json.load(open("gpurun_out/${TAG}_bench_%s.json" % f))
    print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("cpu_baseline", {}).get("value"))
PY
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
