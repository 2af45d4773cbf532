#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_c2.log 2>&1
echo "rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|rank_count_kernel|auroc_kernel" -s 12 -c 4 -o gpurun_out/prof_r01_c2 -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_full_c2.log | cut -c1-200
