#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_c2_n2.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/bench_c2_n2.log | cut -c1-1200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_gpu_check.py > gpurun_out/dist_check.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/dist_check.log
