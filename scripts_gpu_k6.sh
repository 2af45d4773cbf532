#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "auroc or one_vs_best" 2>&1 | tail -4
for cfg in c4 c5; do timeout 600 python tools/bench_fast.py $cfg 2>&1 | tail -8; done
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
