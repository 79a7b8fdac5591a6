# multi-GPU bench lines exactly as the driver launches them: our arm, then the reference arm
cd $GRAFT_REPO_ROOT
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r3_bench_${N}gpu.err | tail -1 > gpurun_out/r3_bench_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2> gpurun_out/r3_bench_reference_${N}gpu.err | tail -1 > gpurun_out/r3_bench_reference_${N}gpu.json
tail -3 gpurun_out/r3_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r3_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('ours n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',d['e2e'] and round(d['e2e']['value'],1), 'h2d GB/s per gpu', d['e2e'] and d['e2e'].get('h2d_gb_s_per_gpu'), d.get('numa'))
r=json.loads(open('gpurun_out/r3_bench_reference_${N}gpu.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('unit'), r.get('cpu_baseline'))
PY
