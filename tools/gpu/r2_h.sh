#!/bin/bash
# round 2, call H: where the supervised path and the e2e call spend their time; sanitizer pass
mkdir -p gpurun_out
python tools/prof_supervised.py 200 > gpurun_out/r2h_sup.log 2>&1
python tools/prof_supervised.py 200 --fast >> gpurun_out/r2h_sup.log 2>&1
cat gpurun_out/r2h_sup.log | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launch_sup.csv python tools/prof_supervised.py 20 > gpurun_out/r2h_ncu_sup.log 2>&1
python tools/summarize_launches.py gpurun_out/r2h_launch_sup.csv | head -24
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launch_supf.csv python tools/prof_supervised.py 20 --fast > gpurun_out/r2h_ncu_supf.log 2>&1
python tools/summarize_launches.py gpurun_out/r2h_launch_supf.csv | head -24
python tools/profile_e2e.py > gpurun_out/r2h_e2e.log 2>&1; head -40 gpurun_out/r2h_e2e.log | cut -c1-200
TMO=300 bash tools/gpu/sanitize.sh
