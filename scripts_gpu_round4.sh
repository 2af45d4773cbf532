#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
TAILN=6 run tests_gpu python -m pytest tests -q -m gpu -x
for b in 8 16 32 64; do
  echo "=== band $b"
  MN_GEMM_BAND=$b timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, d['clocks'])"
done
