"""Robustness probe: K3 single-field latency and 32-field throughput on other map classes (empty, sparse iid, dense
blocks, corridors) at 2048^2 -- the tuning constants were chosen on the cfg3 recipe."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp
from tests.common import block_map

side = 2048
rng = np.random.default_rng(0)
maps = {
    "empty": np.zeros((side, side), np.int8),
    "iid 10%": (rng.random((side, side)) < 0.10).astype(np.int8) * 100,
    "blocks 8x8 30%": block_map(side, side, 0.30, 8, 1),
    "blocks 32x32 35%": block_map(side, side, 0.35, 32, 2),
}
cor = np.zeros((side, side), np.int8)
for y in range(16, side, 32):                       # horizontal walls with alternating gaps: a serpentine
    cor[y:y + 4, :] = 100
    gap = 8 if (y // 32) % 2 == 0 else side - 40
    cor[y:y + 4, gap:gap + 32] = 0
maps["serpentine corridors"] = cor
ctx = hpp.Context(0)
for name, occ in maps.items():
    occ = occ.copy()
    c = side // 2
    occ[c - 2:c + 3, c - 2:c + 3] = 0
    m = hpp.Map(ctx, occ, 0.2)
    for nf in (1, 32):
        goals = np.array([(c - 2 + (i % 5), c - 2 + (i // 5) % 5) for i in range(nf)], np.int32)
        cost = torch.empty((nf, side, side), dtype=torch.float64, device="cuda:0")
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            m.heuristic_batch_dev(hpp.VARIANT_RS, goals, cost); ctx.sync()
            ts.append(time.perf_counter() - t0)
        fin = torch.isfinite(cost[0])
        print(f"{name:22s} fields={nf:3d}: {min(ts) * 1e3 / nf:8.3f} ms/field  reach={int(fin.sum()) / side / side:.2f} "
              f"max={float(cost[0][fin].max()):.0f}")
    m.close()
