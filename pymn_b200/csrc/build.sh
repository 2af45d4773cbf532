#!/bin/bash
# Build libpymn_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -uo pipefail
cd "$(dirname "$0")"
OUT=../libpymn_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
SRCS="lib rank_normalize rank_warp rank_count centroids limbs auroc auroc_count gemm_tcgen05 gene_stats sparse supervised"
mkdir -p build
pids=()
for f in $SRCS; do
  $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
if [ $fail -ne 0 ]; then grep -h -B2 -A6 "error" build/*.log | head -80; echo "build failed"; exit 1; fi
OBJS=""
for f in $SRCS; do OBJS="$OBJS build/$f.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJS || exit 1
echo "built $(realpath $OUT)"
