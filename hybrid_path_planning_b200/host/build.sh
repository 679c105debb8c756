#!/usr/bin/env bash
# Builds libhpp_host.so: the C++ host shells (drop-in classes) on top of libhpp_b200.so.
set -euo pipefail
cd "$(dirname "$0")"
g++ -std=c++17 -O2 -fPIC -Wall -shared -o ../libhpp_host.so shells.cpp -L.. -l:libhpp_b200.so -Wl,-rpath,'$ORIGIN'
echo "built hybrid_path_planning_b200/libhpp_host.so"
g++ -std=c++17 -O2 -Wall -o ../shell_demo shell_demo.cpp -L.. -l:libhpp_host.so -l:libhpp_b200.so -Wl,-rpath,'$ORIGIN'
echo "built hybrid_path_planning_b200/shell_demo"

# Drop-in check (only where the reference checkout is present): the reference's UNMODIFIED planner nodes compiled against
# the shells inside the stand-in ROS of tests/ros_mock; the binaries travel to the GPU box for tests/test_gpu_dropin_node.py.
CXX="${CXX:-g++}"
CXXFLAGS="-std=c++17 -O2 -w"
REF="${HPP_REF:-/root/reference}"
MOCK="$(cd ../../tests/ros_mock && pwd)"
if [ -f "$REF/reeds_shepp_hybrid_astar/src/main_path_planning.cpp" ] && [ -d "$MOCK" ]; then
  for pkg in rs:reeds_shepp_hybrid_astar db:dubins_hybrid_star; do
    tag="${pkg%%:*}"; dir="${pkg##*:}"
    tmp="$(mktemp -d)"
    cp "dropin/$tag/"*.h "$tmp/"
    cp "$REF/$dir/include/main_path_planning.h" "$tmp/"      # temporary, outside the repository
    $CXX $CXXFLAGS -Dmain=reference_node_main -I"$tmp" -I"$MOCK/include" -I. -I../../include \
      -c "$REF/$dir/src/main_path_planning.cpp" -o "$tmp/node.o"
    $CXX $CXXFLAGS -I"$tmp" -I"$MOCK/include" -I. -I../../include "$MOCK/node_harness.cpp" "$tmp/node.o" \
      -L.. -l:libhpp_host.so -l:libhpp_b200.so -Wl,-rpath,'$ORIGIN' -o "../ref_node_$tag"
    rm -rf "$tmp"
    echo "built hybrid_path_planning_b200/ref_node_$tag (reference node source, unmodified, on the shells)"
  done
fi

# map_inflate: the reference's UNMODIFIED node source (it has no header: the harness includes the .cpp) against the cv shim
# of dropin/mi (opencv2/opencv.hpp -> hpp_cv_threshold_otsu / hpp_cv_dilate_ellipse) inside the stand-in ROS
if [ -f "$REF/map_inflate/src/map_inflate_node.cpp" ] && [ -d "$MOCK" ]; then
  $CXX $CXXFLAGS -Dmain=reference_node_main -Idropin/mi -I"$MOCK/include" -I. -I../../include -I"$REF/map_inflate/src" \
    "$MOCK/mi_node_harness.cpp" -L.. -l:libhpp_host.so -l:libhpp_b200.so -Wl,-rpath,'$ORIGIN' -o ../ref_node_mi
  echo "built hybrid_path_planning_b200/ref_node_mi (reference map_inflate node source, unmodified, on the cv shim)"
fi
