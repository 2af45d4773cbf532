#!/bin/bash
# round 2, call V: full test suite, the full bench line on one GPU, ncu of the final K1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2v_tests.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2v_bench.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/r2v_bench.log | tail -1 | python tools/bench_summary.py
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r2v_ref.log | cut -c1-300
ncu --clock-control none --set full --import-source on -k regex:rank_count_kernel -c 2 -o gpurun_out/r2v_k1 -f python tools/bench_k1.py 200000 2000 > gpurun_out/r2v_ncu_k1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
