#!/usr/bin/env bash
# oracle/build.sh -- builds the test-infrastructure libraries (never the product):
#   oracle/liborc.so            the C restatement (always)
#   oracle/_ref/libref_rs.so    the reference's own RS planner TUs   (only if the reference tree is present)
#   oracle/_ref/libref_db.so    the reference's own Dubins planner TUs (idem)
# The reference sources are compiled where they lie ($HPP_REF or /root/reference); nothing is copied.
# Flags: -O2 (upstream's default build is -O0 with asserts; asserts are kept on), x86-64 baseline so
# there is no FMA contraction, -ffp-contract=off to make that explicit.
set -euo pipefail
cd "$(dirname "$0")"
CFLAGS="-O2 -fPIC -ffp-contract=off -fno-fast-math -Wall -Wno-unused-function"
gcc -std=gnu99 $CFLAGS -shared -o liborc.so orc_map.c orc_heur.c orc_rs.c orc_dubins.c orc_planner.c -lm
echo "built oracle/liborc.so"
REF="${HPP_REF:-/root/reference}"
if [ -d "$REF/reeds_shepp_hybrid_astar/src" ]; then
  mkdir -p _ref
  RS="$REF/reeds_shepp_hybrid_astar"; DB="$REF/dubins_hybrid_star"
  CXX="g++ -std=c++17 $CFLAGS -w -fvisibility=hidden -Iref_shim -I. -include ref_shim/ros_std_includes.hpp"
  $CXX -include ref_standin/Grid_map.h -I"$RS/include" -shared -Wl,-Bsymbolic -o _ref/libref_rs.so \
      "$RS/src/HybridAstar.cpp" "$RS/src/CarData.cpp" "$RS/src/reeds_shepp_path.cpp" ref_driver_rs.cpp \
      -L. -l:liborc.so -Wl,-rpath,'$ORIGIN/..'
  $CXX -I"$DB/include" -shared -Wl,-Bsymbolic -o _ref/libref_db.so \
      "$DB/src/HybridAstar.cpp" "$DB/src/Grid_map.cpp" "$DB/src/CarData.cpp" "$DB/src/dubins_path.cpp" \
      "$DB/src/sat_collision_checker.cc" ref_driver_db.cpp
  echo "built oracle/_ref/libref_rs.so and libref_db.so from $REF"
else
  echo "reference tree not present ($REF): keeping prebuilt oracle/_ref/ if any"
fi
