cd $GRAFT_REPO_ROOT
TAG=${1:-o}
(timeout 900 python -m pytest tests/test_ofdm_gpu.py tests/test_golden_gpu.py tests/test_robustness_gpu.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/${TAG}_pytest.log 2>&1
(timeout 600 python bench.py --no-cpu-baseline --no-channel-leg --e2e-steps 10 2>&1 | tail -1) > gpurun_out/${TAG}_bench_ofdm.json 2>&1
(DABGPU_OVERLAP_PLAIN=1 timeout 600 python bench.py --no-cpu-baseline --no-channel-leg --e2e-steps 10 2>&1 | tail -1) > gpurun_out/${TAG}_bench_ofdm_plain.json 2>&1
(DABGPU_NO_OVERLAP=1 timeout 600 python bench.py --no-cpu-baseline --no-channel-leg --e2e-steps 10 2>&1 | tail -1) > gpurun_out/${TAG}_bench_ofdm_serial.json 2>&1
(timeout 600 python bench.py --workload full --steps 60 --no-cpu-baseline --e2e-steps 10 2>&1 | tail -1) > gpurun_out/${TAG}_bench_full.json 2>&1
cat gpurun_out/${TAG}_pytest.log; python - <<PY
import json
for f in ('bench_ofdm','bench_ofdm_plain','bench_ofdm_serial','bench_full'):
    try:
        d=json.loads(open('gpurun_out/${TAG}_'+f+'.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'frames',d['frames_demodulated'])
    except Exception as e:
        print(f,'FAILED',e, open('gpurun_out/${TAG}_'+f+'.json').read()[-600:])
PY
