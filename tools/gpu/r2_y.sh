#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29517 tests/dist_gpu_check.py > gpurun_out/r2y_dist.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r2y_dist.log | cut -c1-300
MN_VOTES_P2P=0 timeout 600 $TR --master-port 29520 tests/dist_gpu_check.py > gpurun_out/r2y_dist_nccl.log 2>&1; echo "dist (NCCL path) rc=$?"; tail -1 gpurun_out/r2y_dist_nccl.log | cut -c1-200
timeout 900 $TR --master-port 29518 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/r2y_bench.log | tail -1 | python tools/bench_summary.py
grep -i "error\|Traceback" gpurun_out/r2y_bench.log | head -5
