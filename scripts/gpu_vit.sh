# Viterbi iteration: channel parity tests + full-chain bench (no ncu)
cd $GRAFT_REPO_ROOT
TAG=${1:-v}
(timeout 900 python -m pytest tests/test_channel_gpu.py tests/test_golden_gpu.py tests/test_adapters_gpu.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 600 python bench.py --workload full --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full.json 2>&1
cat gpurun_out/${TAG}_pytest.log; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_full.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'kernel_ms',d['kernel_ms'],'vit_mbit',d.get('viterbi_mbit_s'))
PY
