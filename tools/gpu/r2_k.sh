#!/bin/bash
# round 2, call K (2 GPUs): all GPU tests (incl. the multi-GPU bit-identity test), K1 micro-benchmark, bench on 2 GPUs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2k_tests.log | cut -c1-300
python tools/bench_k1.py 200000 2000 > gpurun_out/r2k_k1.log 2>&1
python tools/bench_k1.py 500000 3000 >> gpurun_out/r2k_k1.log 2>&1
cat gpurun_out/r2k_k1.log
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2k_bench_n$N.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/r2k_bench_n$N.log | tail -1 | python tools/bench_summary.py
