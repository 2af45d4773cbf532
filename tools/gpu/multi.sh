#!/bin/bash
# N-GPU check: bit-identity of every path, then the bench line (headline + configs block) on N GPUs.
# usage: gpurun --gpus N -- 'NGPU=N bash tools/gpu/multi.sh'
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 tests/dist_gpu_check.py > gpurun_out/dist_check_n$N.log 2>&1
echo "dist_check rc=$?"; tail -3 gpurun_out/dist_check_n$N.log | cut -c1-400
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/bench_n$N.log | tail -1 | python tools/bench_summary.py
