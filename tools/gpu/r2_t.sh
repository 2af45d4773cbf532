#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2t_tests.log | cut -c1-300
timeout 120 python tools/prof_fast.py c4 3 2>&1 | tail -1
MN_DEN_CUDA_CORES=1 timeout 120 python tools/prof_fast.py c4 3 2>&1 | tail -1
timeout 120 python tools/prof_fast.py c1 3 2>&1 | tail -1
