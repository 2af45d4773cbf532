#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_fast.py 2 71429 256 150 > gpurun_out/plain_k6.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:auroc_ss_kernel -s 2 -c 1 -o gpurun_out/prof_k6 -f python tools/bench_fast.py 2 71429 256 150 > gpurun_out/ncu_k6.log 2>&1
echo rc=$?; tail -3 gpurun_out/plain_k6.log
