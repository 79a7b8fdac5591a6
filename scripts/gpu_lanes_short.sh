# quick iteration on the lane Viterbi: channel parity tests + one full-chain bench
cd $GRAFT_REPO_ROOT
TAG=${1:-s}
(timeout 600 python -m pytest tests/test_channel_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 600 python bench.py --workload full --steps 60 --no-cpu-baseline --e2e-steps 4 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full.json 2>&1
cat gpurun_out/${TAG}_pytest.log; python - <<PY
import json
for f in ('bench_full',):
    try:
        d=json.loads(open('gpurun_out/${TAG}_'+f+'.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'kernel_ms/step',{k:round(v/d['steps'],4) for k,v in d['kernel_ms'].items()},'vit_mbit',round(d.get('viterbi_mbit_s',0),1))
    except Exception as e:
        print(f,'FAILED',e, open('gpurun_out/${TAG}_'+f+'.json').read()[-600:])
PY
