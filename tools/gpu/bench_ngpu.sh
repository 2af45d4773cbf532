#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c2_n$N.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_c2_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],1))"
