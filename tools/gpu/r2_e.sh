#!/bin/bash
# round 2, call E: K6 alone -- timings at the c4 / c5 / c2 shapes, one ncu --set full capture of the counting kernel
mkdir -p gpurun_out
python tools/bench_auroc.py 300 7 71429 150 5 gamma > gpurun_out/r2e_k6.log 2>&1
python tools/bench_auroc.py 150 1 1000000 50 3 gamma >> gpurun_out/r2e_k6.log 2>&1
python tools/bench_auroc.py 40 4 50000 10 5 gamma >> gpurun_out/r2e_k6.log 2>&1
python tools/bench_auroc.py 1050 7 71429 150 3 gamma >> gpurun_out/r2e_k6.log 2>&1
cat gpurun_out/r2e_k6.log
ncu --set full --clock-control none --import-source on -k regex:auroc_count_kernel -c 1 -o gpurun_out/r2e_k6 -f python tools/bench_auroc.py 148 7 71429 150 1 gamma > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
