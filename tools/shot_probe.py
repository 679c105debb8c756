"""Probe for K7 (hpp_rs_shot) on a block map: a batch of analytic shots, for ncu instruction counts / timings.
usage: python tools/shot_probe.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp  # noqa: E402
from hybrid_path_planning_b200.capi import NODE_DTYPE  # noqa: E402
from tests.common import block_map, random_poses  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
occ = block_map(400, 300, 0.02, 3, seed=5)
ctx = hpp.Context(0)
m = hpp.Map(ctx, occ, 0.2)
poses = random_poses(4 * n, 400, 300, margin=3.0, seed=11)
free = poses[m.collide_circles(poses) == 0][:2 * n]
k = len(free) // 2
starts = np.zeros(k, NODE_DTYPE)
starts["x"], starts["y"], starts["heading"] = free[:k, 0], free[:k, 1], free[:k, 2]
goals = free[k:2 * k]
for rep in range(3):
    t0 = time.perf_counter()
    shots, samples = m.rs_shot(starts, goals, samples_cap=1 << 22)
    dt = time.perf_counter() - t0
    print(f"rs_shot {k} shots: {dt * 1e3:.2f} ms host call, found {int(shots['found'].sum())}")
