#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
summ() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(f'gpurun_out/{sys.argv[1]}.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],1), d['clocks']['sm_mhz'])
PY
}
TAILN=15 run t_corr python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k corr
grep -q passed gpurun_out/t_corr.log || exit 1
TAILN=6 run t_all python -m pytest tests -q -m gpu -x
run b_pair python bench.py --steps 3 --warmup 3 --no-cpu-baseline >/dev/null; summ b_pair
MN_GEMM_1CTA=1 run b_1cta python bench.py --steps 3 --warmup 3 --no-cpu-baseline >/dev/null; summ b_1cta
MN_SPILL_GB=0 run b_pair_nospill python bench.py --steps 3 --warmup 3 --no-cpu-baseline >/dev/null; summ b_pair_nospill
