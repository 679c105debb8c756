"""Timing probe for K1 (per-cell tables + obstacle/distance layers) on the cfg3 map.  usage: python tools/layers_probe.py [side] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp
from tests.common import cfg3_map
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
import numpy as np
density = sys.argv[3] if len(sys.argv) > 3 else "dense"      # dense = the cfg3 recipe, sparse = 0.1 % obstacle cells
if density == "sparse":
    occ = (np.random.default_rng(1).random((side, side)) < 0.001).astype(np.int8) * 100
elif density == "verysparse":
    occ = (np.random.default_rng(1).random((side, side)) < 0.00001).astype(np.int8) * 100
else:
    occ, _ = cfg3_map(side, seed=1)
ctx = hpp.Context(0)
d_occ = torch.from_numpy(occ).cuda()
m = hpp.Map(ctx, (side, side), 0.2, device_occ=d_occ)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
ts = []
for k in range(reps + 2):
    with torch.cuda.stream(stream):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); m.update_dev(d_occ); b.record(stream)
    ctx.sync(); ts.append(a.elapsed_time(b))
t = min(ts[2:])
print(f"K1 side={side} {density}: {t:.3f} ms  {side * side / t / 1e3:.0f} Mcells/s  {9 * side * side / t / 1e6:.0f} GB/s algorithmic (9 B/cell)")
