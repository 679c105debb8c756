"""Tuning probe for the K3 heuristic kernel (not a test, not the bench): times hpp_heuristic_batch_dev on the cfg3 map
for the current HPP_HEUR_* environment and prints ms per field.  Usage: python tools/heur_tune.py [side] [n_fields] [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp  # noqa: E402
from tests.common import cfg3_map, parking_map  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
occ, goal = cfg3_map(side, seed=1)
if os.environ.get("PROBE_MAP") == "parking":   # the cfg5 map class
    occ = parking_map(side, seed=0)
    fr = np.argwhere(occ == 0)
    goal = (int(fr[len(fr) // 2][1]), int(fr[len(fr) // 2][0]))
ctx = hpp.Context(0)
m = hpp.Map(ctx, occ, 0.2)
rng = np.random.default_rng(0)
goals = [goal]
free = np.argwhere(occ == 0)
for i in range(nf - 1):
    y, x = free[rng.integers(len(free))]
    goals.append((int(x), int(y)))
goals = np.asarray(goals, np.int32)
cost = torch.empty((nf, side, side), dtype=torch.float64, device="cuda:0")
L = hpp.lib()
import ctypes
gp = goals.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
ts = []
for r in range(reps + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = L.hpp_heuristic_batch_dev(m.h, variant, goals.ctypes.data, nf, cost.data_ptr(), None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ctx.check(st)
    ts.append((t1 - t0) * 1e3)
fin = torch.isfinite(cost[0])
print(f"side={side} fields={nf} variant={variant} env={ {k: v for k, v in os.environ.items() if k.startswith('HPP_HEUR')} } "
      f"ms/batch={min(ts[1:]):.3f} ms/field={min(ts[1:]) / nf:.3f} Mcells/s={nf * side * side / min(ts[1:]) / 1e3:.1f} "
      f"reach={int(fin.sum())} sum={float(cost[0][fin].sum()):.3f}")
