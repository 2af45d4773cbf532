#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
summ() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(f'gpurun_out/{sys.argv[1]}.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1], round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],1), d['clocks']['sm_mhz'])
PY
}
TAILN=6 run t_k python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -m gpu -x
run b_a python bench.py --steps 5 --warmup 3 --no-cpu-baseline >/dev/null; summ b_a
MN_SPILL_GB=0 run b_nospill python bench.py --steps 3 --warmup 3 --no-cpu-baseline >/dev/null; summ b_nospill
timeout 1200 ncu --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:"gemm_kernel|vote_spill|rank_count_kernel|auroc_kernel" -s 12 -c 4 --csv --log-file gpurun_out/traffic_c2.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"; tail -3 gpurun_out/ncu_traffic.log | cut -c1-200
