#!/bin/bash
# round 2, call L: GPU tests, then the full bench line on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2l_tests.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/r2l_bench.log | tail -1 | python tools/bench_summary.py
grep '^{' gpurun_out/r2l_bench.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e calls', d['e2e'].get('calls_ms'))"
