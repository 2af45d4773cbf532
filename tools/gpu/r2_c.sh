#!/bin/bash
# round 2, call C: exact int8 votes
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-1500} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-1200; }
TAILN=30 run r2c_votes python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "votes_exact or votes_fast"
TAILN=30 run r2c_headline python -m pytest tests/test_gpu_headline.py -q -m gpu
TAILN=8 run r2c_all python -m pytest tests -q -m gpu
TAILN=1 run r2c_bench python bench.py --steps 5 --warmup 3
