#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-1500; }
TAILN=6 run t_corr python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -m gpu -x
run b_spill python bench.py --steps 3 --warmup 3 --no-cpu-baseline
MN_SPILL_TB=2 run b_spill_tb2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
MN_SPILL_TB=32 run b_spill_tb32 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
MN_SPILL_GB=0 run b_nospill python bench.py --steps 3 --warmup 3 --no-cpu-baseline
