#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29512 tests/dist_gpu_check.py > gpurun_out/dist_check.log 2>&1
echo "dist_check rc=$?"; tail -3 gpurun_out/dist_check.log | cut -c1-300
$TR --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_c2_n$N.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_c2_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],1))"
$TR --master-port 29513 tools/bench_fast.py c4 > gpurun_out/fast_c4_n$N.log 2>&1
echo "fast rc=$?"; tail -6 gpurun_out/fast_c4_n$N.log
