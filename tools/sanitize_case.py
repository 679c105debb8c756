"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): layers, K4 through the filter path, a few
Reeds-Shepp and Dubins queries (pipelined planner with its heap warp, two-pass when HPP_SAN_BIG=1), one heuristic field.
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
(compute-sanitizer is closed on the gpurun pool of this build -- "runs under it have left GPUs needing a reset" -- so this
was only ever run plain there; kept for a box where the tool is available.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp  # noqa: E402
from tests.common import block_map, parking_map, random_poses  # noqa: E402

ctx = hpp.Context(0)
occ = parking_map(128, seed=2)
m = hpp.Map(ctx, occ, 0.2)
poses = random_poses(20000, 128, 128, margin=-1.0, seed=1)
v = m.collide_circles(poses)                       # >= 16 384 poses: filter + list kernels
free = poses[(v == 0)]
free = free[(np.abs(free[:, 0]) < 9) & (np.abs(free[:, 1]) < 9)]
nq = 4800 if os.environ.get("HPP_SAN_BIG") else 12
starts, goals = free[:nq], free[nq:2 * nq]
assert len(goals) == nq
plans, traj = m.plan_batch(hpp.VARIANT_RS, starts, goals)
print("rs: solved", int((plans["n_states"] > 1).sum()), "of", nq, "iterations", int(plans["iterations"].sum()))
plans, traj = m.plan_batch(hpp.VARIANT_DUBINS, starts[:6], goals[:6])
print("dubins: solved", int((plans["n_states"] > 1).sum()), "iterations", int(plans["iterations"].sum()))
h = m.heuristic_rs((64, 64), (0, 0))
print("field finite", int(np.isfinite(h[0] if isinstance(h, tuple) else h).sum()), "colliding", float(v.mean()))
m.close()
