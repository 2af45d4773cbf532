#!/bin/bash
# round 2, call I: GPU tests after the cell-sharded centroid path / fixed-point centroids / supervised host work
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2i_tests.log | cut -c1-300
python tools/prof_supervised.py 300 > gpurun_out/r2i_sup.log 2>&1
python tools/prof_supervised.py 300 --fast >> gpurun_out/r2i_sup.log 2>&1
tail -3 gpurun_out/r2i_sup.log
python tools/prof_fast.py c4 3 > gpurun_out/r2i_c4.log 2>&1; tail -3 gpurun_out/r2i_c4.log
python tools/prof_fast.py c1 3 > gpurun_out/r2i_c1.log 2>&1; tail -2 gpurun_out/r2i_c1.log
