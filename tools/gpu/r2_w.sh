#!/bin/bash
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for i in 1 2; do
timeout 600 $TR --master-port 2951$i bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench_$i.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/r2w_bench_$i.log | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['steps_ms'], d['e2e']['calls_ms'], d['clocks']); c=d['configs']
for k in ('c4','c5'): print(k, round(c[k]['ms_per_step'],2), {a: round(b,2) for a,b in c[k]['stage_ms'].items()})"
done
