#!/usr/bin/env bash
# Builds libhpp_b200.so (the product) for sm_100a only, in-tree.
#   -fmad=false : the reference build (baseline x86-64) never contracts a*b+c; neither may we (bit parity)
#   -lineinfo   : ncu source pages
set -euo pipefail
cd "$(dirname "$0")"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 ${HPP_NVCC_EXTRA:-}"
SRCS="map.cu edt.cu inflate.cu collide.cu heuristic.cu expand.cu shots.cu planner.cu mgpu.cu"
mkdir -p build
pids=()
for f in $SRCS; do
  o="build/${f%.cu}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ] || [ ../../include/hpp_b200.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=$(for f in $SRCS; do echo "build/${f%.cu}.o"; done)
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libhpp_b200.so $OBJS -lcudart -ldl
echo "built hybrid_path_planning_b200/libhpp_b200.so"
