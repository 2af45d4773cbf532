#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
TAILN=25 run tests_gpu python -m pytest tests -q -m gpu -x
TAILN=3 run bench_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
