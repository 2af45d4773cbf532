#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
run bench_c2 python bench.py --steps 5 --warmup 3
TAILN=8 run fast_c1 python tools/bench_fast.py c1
TAILN=8 run fast_c4 python tools/bench_fast.py c4
TAILN=8 run fast_c5 python tools/bench_fast.py c5
run sup_default python tools/bench_supervised.py 200
run sup_fast python tools/bench_supervised.py 500 --fast
