#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config c2_small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_small.csv \
   python bench.py --config c2_small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|rank_count|auroc_kernel" -s 4 -c 4 -o gpurun_out/prof_r01_c2_small -f \
   python bench.py --config c2_small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200
