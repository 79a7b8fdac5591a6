# lane-per-trellis Viterbi iteration: channel parity tests (auto / lanes / warps) + full-chain bench with each mapping
cd $GRAFT_REPO_ROOT
TAG=${1:-l}
(timeout 900 python -m pytest tests/test_channel_gpu.py tests/test_golden_gpu.py tests/test_adapters_gpu.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 600 python bench.py --workload full --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full.json 2>&1
(DABGPU_VIT_LANES_PREP=1 timeout 600 python bench.py --workload full --steps 40 --no-cpu-baseline --e2e-steps 4 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full_prep.json 2>&1
(timeout 600 python bench.py --workload full --streams 512 --steps 50 --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full_512.json 2>&1
cat gpurun_out/${TAG}_pytest.log; python - <<PY
import json
for f in ('bench_full','bench_full_prep','bench_full_512'):
    try:
        d=json.loads(open('gpurun_out/${TAG}_'+f+'.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms',{k:round(v,2) for k,v in d['kernel_ms'].items()},'vit_mbit',round(d.get('viterbi_mbit_s',0),1),'launches',d['gpu_launches'],d['counters'])
    except Exception as e:
        print(f,'FAILED',e, open('gpurun_out/${TAG}_'+f+'.json').read()[-600:])
PY
