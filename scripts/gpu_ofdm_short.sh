cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_ofdm_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -3)
timeout 300 python bench.py --no-cpu-baseline --no-channel-leg --steps 100 --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step',round(d['ms_per_step'],4),'value',round(d['value'],1),'demod',round(d['kernel_ms']['ofdm_demod']/d['steps'],4),'ctl',round(d['kernel_ms']['ofdm_ctl']/d['steps'],4))"
