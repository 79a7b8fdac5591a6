cd $GRAFT_REPO_ROOT
TAG=${1:-dp}
(timeout 900 python -m pytest tests/test_channel_gpu.py tests/test_golden_gpu.py tests/test_adapters_gpu.py tests/test_autoconfig_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 600 python bench.py --workload full --steps 60 --no-cpu-baseline --e2e-steps 20 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full.json 2>&1
(DABGPU_CHAN_INLINE=1 timeout 600 python bench.py --workload full --steps 60 --no-cpu-baseline --e2e-steps 20 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full_inline.json 2>&1
cat gpurun_out/${TAG}_pytest.log; python - <<PY
import json
for f in ('bench_full','bench_full_inline'):
    try:
        d=json.loads(open('gpurun_out/${TAG}_'+f+'.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'vit',round(d['viterbi_mbit_s'],1), {k:d['counters'][k] for k in ('superframes_ok','au_ok','fibs_crc_ok')})
    except Exception as e:
        print(f,'FAILED',e, open('gpurun_out/${TAG}_'+f+'.json').read()[-600:])
PY
