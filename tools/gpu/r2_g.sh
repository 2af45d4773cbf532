#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:auroc_count_kernel -c 1 -o gpurun_out/r2g_k6 -f python tools/bench_auroc.py 148 7 71429 150 1 gamma > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
