#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
summ() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(f'gpurun_out/{sys.argv[1]}.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],1), d['clocks']['sm_mhz'])
PY
}
for w in 24 20 16 24 20; do
MN_SPILL_WARPS=$w run b_w$w python bench.py --steps 5 --warmup 3 --no-cpu-baseline >/dev/null; summ b_w$w
done
