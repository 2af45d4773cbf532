#!/bin/bash
# first GPU bring-up: each group in its own process with a timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/$name.log; }
run k_rank python -m pytest tests/test_gpu_kernels.py -q -x -k "rank_normalize or drop_mode or centroids"
run k_auroc python -m pytest tests/test_gpu_kernels.py -q -k "auroc_kernel or one_vs_best"
run k_gemm python -m pytest tests/test_gpu_kernels.py -q -k "votes_fast"
run k_corr python -m pytest tests/test_gpu_kernels.py -q -k "corr_hist"
run e2e_fast python -m pytest tests/test_gpu_e2e.py -q -k "fast or train_model or pretrained"
run e2e_rest python -m pytest tests/test_gpu_e2e.py -q -k "not (fast or train_model or pretrained)"
