#!/bin/bash
one() { echo "=== $*"; env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, d['clocks']['sm_mhz'])"; }
one MN_GEMM_BAND=8
one MN_GEMM_BAND=16
one MN_GEMM_BAND=24
one MN_GEMM_BAND=32
one MN_GEMM_BAND=48
one MN_GEMM_BAND=64
