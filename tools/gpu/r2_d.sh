#!/bin/bash
# round 2, call D: where the centroid path's time goes (launch lists under ncu; plain runs first)
mkdir -p gpurun_out
python tools/prof_fast.py c1 4 > gpurun_out/r2d_c1.log 2>&1; tail -4 gpurun_out/r2d_c1.log
python tools/prof_fast.py c4 3 > gpurun_out/r2d_c4.log 2>&1; tail -3 gpurun_out/r2d_c4.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launch_c1.csv python tools/prof_fast.py c1 2 > gpurun_out/r2d_ncu_c1.log 2>&1
python tools/summarize_launches.py gpurun_out/r2d_launch_c1.csv | head -30
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launch_c4.csv python tools/prof_fast.py c4 1 > gpurun_out/r2d_ncu_c4.log 2>&1
python tools/summarize_launches.py gpurun_out/r2d_launch_c4.csv | head -30
