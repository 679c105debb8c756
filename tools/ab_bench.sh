#!/usr/bin/env bash
# A/B of two library builds (ab/libA.so, ab/libB.so: tools/ab_build.sh) through bench.py's own e2e and cfg5 sections.
for rep in 1 2; do for v in A B; do
  HPP_B200_LIB=ab/lib$v.so python bench.py --steps 3 --warmup 3 --no-extras --quick 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'cfg5 %.0f q/s %.1f ms' % (d['sharded']['cfg5']['queries_per_s'], d['sharded']['cfg5']['ms']))"
done; done
