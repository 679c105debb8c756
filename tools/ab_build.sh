#!/usr/bin/env bash
# Builds the library twice for an A/B measurement inside one gpurun session: ab/libA.so with $1 as extra nvcc flags,
# ab/libB.so with $2; the in-tree libhpp_b200.so is rebuilt with the default flags afterwards.
#   tools/ab_build.sh "-DHPP_LIBM_INLINE" ""  &&  gpurun -- 'HPP_B200_LIB=ab/libA.so python tools/cfg5_probe.py; HPP_B200_LIB=ab/libB.so ...'
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p ab
for v in A B; do
  if [ $v = A ]; then extra="$1"; else extra="$2"; fi
  rm -rf hybrid_path_planning_b200/csrc/build
  HPP_NVCC_EXTRA="$extra" bash hybrid_path_planning_b200/csrc/build.sh >/dev/null
  cp hybrid_path_planning_b200/libhpp_b200.so ab/lib$v.so
done
rm -rf hybrid_path_planning_b200/csrc/build
bash hybrid_path_planning_b200/csrc/build.sh
