cd $GRAFT_REPO_ROOT
for spc in 0 7 9 11 13 15 19 25 38; do
  if [ $spc -eq 0 ]; then unset DABGPU_DEMOD_SPC; else export DABGPU_DEMOD_SPC=$spc; fi
  timeout 300 python bench.py --no-cpu-baseline --no-channel-leg --steps 60 --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('spc',$spc,'ms/step',round(d['ms_per_step'],4),'demod',round(d['kernel_ms']['ofdm_demod']/d['steps'],4),'ctl',round(d['kernel_ms']['ofdm_ctl']/d['steps'],4))"
done
