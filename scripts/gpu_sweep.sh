cd $GRAFT_REPO_ROOT
for spc in 0 10 13 15 19 25 38; do
  if [ $spc = 0 ]; then unset DABGPU_DEMOD_SPC; else export DABGPU_DEMOD_SPC=$spc; fi
  python bench.py --no-cpu-baseline --steps 60 --e2e-steps 4 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('spc=$spc', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'demod_ms', round(d['kernel_ms']['ofdm_demod']/d['steps'],4), 'ctl_ms', round(d['kernel_ms']['ofdm_ctl']/d['steps'],4))"
done
unset DABGPU_DEMOD_SPC
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
