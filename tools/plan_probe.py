"""Timing probe for hpp_plan_batch on the cfg5 workload (not a test, not the bench)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp
from tests.common import parking_map, random_poses

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = hpp.Context(0)
pm = parking_map(256, seed=0)
cand = random_poses(20000, 256, 256, margin=3.0, seed=100)
m = hpp.Map(ctx, pm, 0.2)
free = cand[:6000][m.collide_circles(cand[:6000]) == 0]      # device verdicts
reps = (2 * npairs + len(free) - 1) // len(free)
free = np.concatenate([free] * max(1, reps))[:2 * npairs]
starts, goals = free[:npairs], free[npairs:2 * npairs]
m.plan_batch(hpp.VARIANT_RS, starts[:8], goals[:8])
for rep in range(2):
    t0 = time.perf_counter(); plans, _ = m.plan_batch(hpp.VARIANT_RS, starts, goals); t1 = time.perf_counter()
    print(f"plan_batch n={npairs}: {1e3*(t1-t0):.1f} ms  -> {npairs/(t1-t0):.0f} q/s  solved={(plans['n_states']>1).sum()} "
          f"iters mean={plans['iterations'].mean():.0f} max={plans['iterations'].max()} via_shot={plans['via_shot'].sum()}")
# heuristic fields alone
cells = np.stack([((goals[:, 0] - m.ox) / m.res).astype(np.int32), ((goals[:, 1] - m.oy) / m.res).astype(np.int32)], 1)
cost = torch.empty((npairs, 256, 256), dtype=torch.float64, device="cuda:0")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.heuristic_batch_dev(hpp.VARIANT_RS, cells, cost); ctx.sync(); t1 = time.perf_counter()
    print(f"heuristic fields n={npairs}: {1e3*(t1-t0):.1f} ms")
