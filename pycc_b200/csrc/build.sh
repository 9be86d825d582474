#!/bin/bash
# Build libb200cc.so for sm_100a (cross-compiles without a GPU).  Usage: build.sh [extra nvcc flags]
# Each translation unit is compiled in its own nvcc process (in parallel), objects are kept under csrc/build/ and
# only rebuilt when their source (or a header) is newer.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libb200cc.so"
OBJ="$HERE/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default "$@")
mkdir -p "$OBJ"
SIG="$(echo "${FLAGS[*]}" | md5sum | cut -c1-8)"
pids=()
objs=()
for src in "$HERE"/*.cu; do
  name="$(basename "$src" .cu)"
  obj="$OBJ/$name.$SIG.o"
  objs+=("$obj")
  stale=0
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then stale=1; fi
  for h in "$HERE"/*.cuh "$HERE/../../include"/*.h; do
    if [ -f "$h" ] && [ "$h" -nt "$obj" ]; then stale=1; fi
  done
  if [ "$stale" = 1 ]; then
    "$NVCC" "${FLAGS[@]}" -c -o "$obj" "$src" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "${objs[@]}"
echo "built $OUT"
