# multi-GPU bench line of our arm only, as the driver launches it
cd $GRAFT_REPO_ROOT
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-spot-check 2> gpurun_out/r2_bench_${N}gpu.err | tail -1 > gpurun_out/r2_bench_${N}gpu.json
tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('ours n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',d['e2e'] and round(d['e2e']['value'],1), 'h2d GB/s per gpu', d['e2e'] and d['e2e'].get('h2d_gb_s_per_gpu'), d.get('numa'))
PY
