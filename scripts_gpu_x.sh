#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-300; }
summ() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(f'gpurun_out/{sys.argv[1]}.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],1), d['clocks']['sm_mhz'])
PY
}
for w in 16 8 12; do
MN_SPILL_WARPS=$w run b_w$w python bench.py --steps 3 --warmup 3 --no-cpu-baseline >/dev/null; summ b_w$w
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vote_spill" -s 1 -c 1 -o gpurun_out/prof_spill_mid -f \
   python bench.py --config c2_mid --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_spill_mid.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_spill_mid.log | cut -c1-200
