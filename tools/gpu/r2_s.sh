#!/bin/bash
# round 2, call S: int8 vote GEMM with k-block-major staging, single CTA and CTA pairs
mkdir -p gpurun_out
for c in 1 2; do
  echo "== MN_I8_CTAS=$c"
  MN_I8_CTAS=$c timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "votes_exact or centroids" 2>&1 | tail -3 | cut -c1-200
  MN_I8_CTAS=$c timeout 300 python -m pytest tests/test_gpu_headline.py tests/test_gpu_e2e.py -x -q -m gpu -k "config4 or config5 or config1 or golden_unsup or golden_pretr or golden_superv or g3000" 2>&1 | tail -3 | cut -c1-200
  MN_I8_CTAS=$c timeout 120 python tools/prof_fast.py c4 3 2>&1 | tail -1
done
