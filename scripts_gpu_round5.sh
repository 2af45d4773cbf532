#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
TAILN=25 run tests_gpu python -m pytest tests -q -m gpu -x
for i in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c2_$i.log
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_c2_$i.log').read()); print('ms/step', round(d['ms_per_step'],1), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['clocks'], 'e2e ms', round(d['e2e']['ms_per_step'],1))"
done
