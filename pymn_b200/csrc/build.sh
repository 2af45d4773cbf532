#!/bin/bash
# Build libpymn_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libpymn_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
pids=()
for f in lib rank_normalize centroids auroc gemm_tcgen05; do
  $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
if [ $fail -ne 0 ]; then cat build/*.log; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/lib.o build/rank_normalize.o build/centroids.o build/auroc.o build/gemm_tcgen05.o
echo "built $(realpath $OUT)"
