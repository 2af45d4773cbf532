#!/bin/bash
# round 2, call J: K1 limb output + centroid path after it
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2j_tests.log | cut -c1-300
python tools/bench_k1.py 200000 2000 > gpurun_out/r2j_k1.log 2>&1
python tools/bench_k1.py 500000 3000 >> gpurun_out/r2j_k1.log 2>&1
cat gpurun_out/r2j_k1.log
python tools/prof_fast.py c4 3 > gpurun_out/r2j_c4.log 2>&1; tail -2 gpurun_out/r2j_c4.log
python tools/prof_fast.py c1 3 > gpurun_out/r2j_c1.log 2>&1; tail -2 gpurun_out/r2j_c1.log
python tools/prof_supervised.py 300 --fast > gpurun_out/r2j_sup.log 2>&1; tail -1 gpurun_out/r2j_sup.log
