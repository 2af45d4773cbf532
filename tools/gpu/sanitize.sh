#!/bin/bash
# compute-sanitizer over the small kernel-level parity tests (race / memory checking).
# memcheck: every kernel family once; racecheck: shared-memory hazards of the kernels that synchronise through
# __syncthreads / __syncwarp (the tcgen05 kernels hand data over through mbarriers and the async proxy, which
# racecheck does not model: they run under memcheck only).  usage: gpurun -- bash tools/gpu/sanitize.sh
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SMALL='test_rank_normalize_bit_exact or test_auroc_kernel_matches_oracle or test_one_vs_best or test_centroids_and_group_sum or test_votes_fast_tcgen05_vs_reference or test_corr_hist_and_votes or test_votes_exact or test_corr_votes_heavy'
timeout ${TMO:-900} $SAN --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$SMALL" -p no:cacheprovider > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitize_memcheck.log | tail -8
RACE='test_rank_normalize_bit_exact or test_auroc_kernel_matches_oracle or test_one_vs_best or test_centroids_and_group_sum'
timeout ${TMO:-900} $SAN --tool racecheck --racecheck-report hazard --error-exitcode 9 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$RACE" -p no:cacheprovider > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | tail -8
