#!/bin/bash
# round 2, call A: full GPU test-suite, smoke, bench (with the configs block)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${TMO:-1500} "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
TAILN=15 run r2a_tests python -m pytest tests -q -m gpu -x --durations=8
run r2a_smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=1 run r2a_bench python bench.py --steps 5 --warmup 3
