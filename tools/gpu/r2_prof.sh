#!/bin/bash
# round 2 profiles: launch list of the bench command, ncu --set full of K1 / K6 / pass 2, full-size counters of
# pass 1 (application replay), all AFTER the plain runs exited 0.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2q_bench_plain.log 2>&1; echo "plain bench rc=$?"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2q_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2q_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r2q_launches_c2.csv > gpurun_out/r2q_launches_c2.txt; head -14 gpurun_out/r2q_launches_c2.txt
python tools/bench_k1.py 200000 2000 > gpurun_out/r2q_k1.log 2>&1 && \
$NCU --set full --import-source on -k regex:rank_count_kernel -c 2 -o gpurun_out/r2q_k1 -f python tools/bench_k1.py 200000 2000 > gpurun_out/r2q_ncu_k1.log 2>&1
python tools/bench_auroc.py 1050 7 71429 150 1 gamma > gpurun_out/r2q_k6.log 2>&1 && \
$NCU --set full --import-source on -k regex:auroc_count_kernel -c 1 -o gpurun_out/r2q_k6_c4 -f python tools/bench_auroc.py 1050 7 71429 150 1 gamma > gpurun_out/r2q_ncu_k6.log 2>&1
python tools/prof_c2.py 50000 1 > gpurun_out/r2q_c2_plain.log 2>&1; cat gpurun_out/r2q_c2_plain.log | tail -1
timeout 900 $NCU --replay-mode application --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats \
  -k regex:'gemm_kernel|vote_spill_kernel' -o gpurun_out/r2q_c2_full -f python tools/prof_c2.py 50000 1 > gpurun_out/r2q_ncu_c2.log 2>&1
echo "app replay rc=$?"; tail -3 gpurun_out/r2q_ncu_c2.log
ls -la gpurun_out/*.ncu-rep | tail -5
