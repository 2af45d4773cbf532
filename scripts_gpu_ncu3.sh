#!/bin/bash
M="gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['stage_ms'].items()}, d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['ms_per_step'],1))" || exit 1
timeout 900 ncu --metrics $M --clock-control none -k regex:"gemm_kernel" -s 2 -c 2 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -E "gemm_kernel" | sed 's/CUtensorMap_st[^)]*)//' | rev | cut -c1-80 | rev
