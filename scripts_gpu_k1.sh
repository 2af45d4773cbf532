#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "rank or drop" 2>&1 | tail -5
python tools/bench_k1.py 200000 2000
MN_RANK_NOCOUNT=1 python tools/bench_k1.py 200000 2000
python tools/bench_k1.py 100000 3000
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
