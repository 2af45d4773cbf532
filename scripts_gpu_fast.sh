#!/bin/bash
mkdir -p gpurun_out
for cfg in c1 c4 c5; do timeout 600 python tools/bench_fast.py $cfg 2>&1 | tail -9; done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c2_a.log
python -c "
import json
d=json.loads(open('gpurun_out/bench_c2_a.log').read()); print('c2 ms/step', round(d['ms_per_step'],1), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['clocks'], 'e2e ms', round(d['e2e']['ms_per_step'],1))"
