# multi-GPU bench exactly as the driver launches it
cd $GRAFT_REPO_ROOT
N=${1:-2}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 4 2>&1 | tail -2 > gpurun_out/multi_ofdm_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 4 --workload full 2>&1 | tail -2 > gpurun_out/multi_full_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 3 2>&1 | tail -2 > gpurun_out/multi_ref_$N.log
cat gpurun_out/multi_ofdm_$N.log gpurun_out/multi_full_$N.log gpurun_out/multi_ref_$N.log
