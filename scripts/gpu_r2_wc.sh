# experiment: write-combined pinned host memory for the H2D side of the end-to-end leg, N GPUs
cd $GRAFT_REPO_ROOT
N=${1:-2}
for wc in "" "--wc-host"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 5 --no-spot-check --no-ofdm-leg $wc 2>/dev/null | tail -1 > gpurun_out/wc_tmp.json
python - <<PY
import json
d=json.loads(open('gpurun_out/wc_tmp.json').read().strip().splitlines()[-1])
print('wc="$wc" n_gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), 'h2d GB/s per gpu', round(d['e2e'].get('h2d_gb_s_per_gpu'),2), d['e2e'].get('host_memory'))
PY
done
