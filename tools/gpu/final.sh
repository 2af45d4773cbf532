#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=4 run tests_gpu python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
run bench_c2 python bench.py --steps 5 --warmup 3
TAILN=8 run fast_c1 python tools/bench_fast.py c1
TAILN=8 run fast_c4 python tools/bench_fast.py c4
TAILN=8 run fast_c5 python tools/bench_fast.py c5
run sup_default python tools/bench_supervised.py 200
run sup_fast python tools/bench_supervised.py 500 --fast
run k1 python tools/bench_k1.py 200000 2000
echo "=== ncu"
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_c2.log 2>&1
echo "rc=$?"
timeout 600 python bench.py --config c2_mid --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_mid.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|rank_count_kernel|auroc_kernel|vote_spill" -s 15 -c 5 -o gpurun_out/prof_r01_mid -f \
   python bench.py --config c2_mid --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_mid.log 2>&1
echo "rc=$?"
