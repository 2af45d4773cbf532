#!/bin/bash
mkdir -p gpurun_out
one() { echo "=== $*"; env "$@" MN_GEMM_BAND=64 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, d['clocks']['sm_mhz'])"; }
one MN_NBINS=16384 MN_GEMM_BN256=1
one MN_NBINS=16384 MN_GEMM_BN256=1 MN_GEMM_STAGES=2
MN_NBINS=16384 MN_GEMM_BN256=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k corr_hist 2>&1 | tail -3
MN_NBINS=16384 MN_GEMM_BN256=1 timeout 600 python -m pytest tests/test_gpu_e2e.py -q -k "default" 2>&1 | tail -3
