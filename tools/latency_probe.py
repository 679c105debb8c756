"""Single-query latency breakdown on the reference's fixture maps (not a test, not the bench).
usage: python tools/latency_probe.py"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp
from tests.common import random_poses

maps = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "maps.npz"))
ctx = hpp.Context(0)
for name, margin in (("map_parking", 0.3), ("gridmap", 10.0)):
    occ = maps[name]
    H, W = occ.shape
    m = hpp.Map(ctx, occ, 0.2)
    cand = random_poses(4000, W, H, res=0.2, margin=margin, seed=11)
    free = cand[m.collide_circles(cand) == 0][:16]
    s, g = free[:8], free[8:16]
    cells = np.stack([((g[:, 0] - m.ox) / m.res).astype(np.int32), ((g[:, 1] - m.oy) / m.res).astype(np.int32)], 1)
    cost = torch.empty((1, H, W), dtype=torch.float64, device="cuda:0")
    m.heuristic_batch_dev(hpp.VARIANT_RS, cells[:1], cost); ctx.sync()
    th = []
    for i in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.heuristic_batch_dev(hpp.VARIANT_RS, cells[i:i + 1], cost); ctx.sync()
        th.append(time.perf_counter() - t0)
    m.plan_batch(hpp.VARIANT_RS, s[:1], g[:1], traj_cap=1 << 16)
    tp, it = [], []
    for i in range(8):
        t0 = time.perf_counter()
        plans, _ = m.plan_batch(hpp.VARIANT_RS, s[i:i + 1], g[i:i + 1], traj_cap=1 << 16)
        tp.append(time.perf_counter() - t0); it.append(int(plans["iterations"][0]))
    tc = []
    for i in range(8):
        t0 = time.perf_counter(); m.collide_circles(s[i:i + 1]); tc.append(time.perf_counter() - t0)
    print(f"{name} {W}x{H}: heuristic field {1e3*np.mean(th):.3f} ms (min {1e3*min(th):.3f}); plan_batch(1) {1e3*np.mean(tp):.3f} ms "
          f"(min {1e3*min(tp):.3f}, iterations {it}); collide_circles(1) {1e3*np.mean(tc):.3f} ms")
    m.close()
