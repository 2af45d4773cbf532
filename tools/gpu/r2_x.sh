#!/bin/bash
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29517 tests/dist_gpu_check.py > gpurun_out/r2x_dist.log 2>&1; echo "dist rc=$?"; tail -1 gpurun_out/r2x_dist.log | cut -c1-200
MN_BENCH_TRACE=6 timeout 600 $TR --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/r2x_bench.log | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['steps_ms'], d['e2e']['calls_ms'], d['clocks']); c=d['configs']
for k in ('c4','c5'): print(k, round(c[k]['ms_per_step'],2), {a: round(b,2) for a,b in c[k]['stage_ms'].items()})"
grep -A16 "traced e2e call" gpurun_out/r2x_bench.log | cut -c1-120
