#!/bin/bash
# round 2, call B: counting AUROC kernel -- kernel tests, headline tests, bench
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-1500} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-900; }
TAILN=25 run r2b_auroc python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "auroc or one_vs_best" --durations=5
TAILN=40 run r2b_headline python -m pytest tests/test_gpu_headline.py -q -m gpu
TAILN=6 run r2b_all python -m pytest tests -q -m gpu
TAILN=1 run r2b_bench python bench.py --steps 5 --warmup 3
