#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-1500} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-900; }
TAILN=12 run r2f_auroc python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "auroc or one_vs_best" --durations=3
python tools/bench_auroc.py 300 7 71429 150 5 gamma > gpurun_out/r2f_k6.log 2>&1
python tools/bench_auroc.py 150 1 1000000 50 3 gamma >> gpurun_out/r2f_k6.log 2>&1
python tools/bench_auroc.py 40 4 50000 10 5 gamma >> gpurun_out/r2f_k6.log 2>&1
python tools/bench_auroc.py 1050 7 71429 150 3 gamma >> gpurun_out/r2f_k6.log 2>&1
python tools/bench_auroc.py 1050 7 71429 150 3 normal >> gpurun_out/r2f_k6.log 2>&1
cat gpurun_out/r2f_k6.log
