#!/usr/bin/env bash
# SASS instruction count per kernel of one object file: tools/sass_size.sh build/planner.o
set -euo pipefail
d=$(mktemp -d)
( cd "$d" && cuobjdump -xelf all "$(realpath "$1")" >/dev/null )
for f in "$d"/*.cubin; do
  nvdisasm -c "$f" | awk '/^\/\/-+ \.text\./{name=$2} /^ +\/\*[0-9a-f]+\*\//{cnt[name]++} END{for(n in cnt) print cnt[n], substr(n,7,90)}'
done | sort -rn
