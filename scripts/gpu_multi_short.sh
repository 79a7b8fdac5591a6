# multi-GPU bench as the driver launches it (our arm only)
cd $GRAFT_REPO_ROOT
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 4 --no-channel-leg 2>&1 | tail -1 > gpurun_out/multi_ofdm_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 4 --workload full 2>&1 | tail -1 > gpurun_out/multi_full_$N.log
python - <<PY
import json
for f in ('ofdm','full'):
    d=json.loads(open('gpurun_out/multi_%s_$N.log'%f).read().strip().splitlines()[-1])
    print(f,'n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1))
PY
