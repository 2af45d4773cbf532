#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
run tests_gpu python -m pytest tests -q -m gpu -s
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_small python bench.py --config c2_small --steps 2 --warmup 1
TAILN=3 run bench_c2 python bench.py --steps 3 --warmup 3
TAILN=3 run bench_ref python bench.py --impl reference --steps 1 --warmup 0
echo "=== ncu launch list"
timeout 600 python bench.py --config c2_small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_small.csv \
   python bench.py --config c2_small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_launch.log
echo "=== ncu full (gemm kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 4 -o gpurun_out/prof_gemm_c2_small \
   python bench.py --config c2_small --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
