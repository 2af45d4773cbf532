#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config c2_small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 || { tail -5 gpurun_out/plain_small.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vote_spill|gemm_kernel" -s 2 -c 2 -o gpurun_out/prof_spill_small -f \
   python bench.py --config c2_small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_spill.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_spill.log
