"""Probe for the product end-to-end path on the cfg3 map: hpp_map_update (host occupancy) + hpp_plan_batch of F
queries with F distinct goal cells (fields stay on the device), timed per phase with HPP_PLAN_TIMING."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp  # noqa: E402
from bench import patch_goals, RES  # noqa: E402
from tests.common import cfg3_map, random_poses  # noqa: E402

side, F = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 148
occ, _ = cfg3_map(side, seed=1)
ctx = hpp.Context(0)
m = hpp.Map(ctx, occ, RES)
res = float(np.float32(RES))
ox = oy = -side * res / 2.0
cells = np.array(patch_goals(side, F))
cand = random_poses(40000, side, side, margin=20.0, seed=5)
free = cand[m.collide_circles(cand) == 0]
print("free poses", len(free), "of", len(cand))
# goals: free poses with distinct cells; starts: free poses
gcell = ((free[:, 0] - ox) / res).astype(int) + side * ((free[:, 1] - oy) / res).astype(int)
_, first = np.unique(gcell, return_index=True)
free_g = free[np.sort(first)]
goals = free_g[:F]
starts = free_g[F:2 * F]
for mode in ("free-pose goals", "patch-cell goals"):
    if mode == "patch-cell goals":
        goals = np.stack([ox + (cells[:, 0] + 0.5) * res, oy + (cells[:, 1] + 0.5) * res, np.zeros(F)], 1)
    for rep in range(3):
        t0 = time.perf_counter()
        m.update(occ)
        t1 = time.perf_counter()
        plans, traj = m.plan_batch(hpp.VARIANT_RS, starts, goals, traj_cap=1 << 20)
        t2 = time.perf_counter()
        print(f"{mode}: update {1e3 * (t1 - t0):.1f} ms, plan_batch {1e3 * (t2 - t1):.1f} ms, solved "
              f"{(plans['n_states'] > 1).sum()}/{F}, iterations mean {plans['iterations'].mean():.0f} max "
              f"{plans['iterations'].max()}, e2e {F * side * side / (t2 - t0) / 1e6:.0f} Mcells/s")
m.close()
