#!/usr/bin/env bash
# Builds libhpp_host.so: the C++ host shells (drop-in classes) on top of libhpp_b200.so.
set -euo pipefail
cd "$(dirname "$0")"
g++ -std=c++17 -O2 -fPIC -Wall -shared -o ../libhpp_host.so shells.cpp -L.. -l:libhpp_b200.so -Wl,-rpath,'$ORIGIN'
echo "built hybrid_path_planning_b200/libhpp_host.so"
g++ -std=c++17 -O2 -Wall -o ../shell_demo shell_demo.cpp -L.. -l:libhpp_host.so -l:libhpp_b200.so -Wl,-rpath,'$ORIGIN'
echo "built hybrid_path_planning_b200/shell_demo"
