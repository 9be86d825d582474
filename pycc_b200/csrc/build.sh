#!/bin/bash
# Build libb200cc.so for sm_100a (cross-compiles without a GPU).  Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libb200cc.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xcompiler -fPIC -Xcompiler -fvisibility=default -shared \
  -o "$OUT" "$HERE/gemm.cu" "$HERE/permute.cu" "$HERE/elementwise.cu" "$HERE/triples.cu" "$HERE/mixed.cu" "$@"
echo "built $OUT"
