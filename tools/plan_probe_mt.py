"""cfg5 throughput probe: nmaps parking maps x npairs queries, one context per map, maps planned concurrently from host
threads.  usage: python tools/plan_probe_mt.py [nmaps] [npairs]"""
import os, sys, time, threading
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp
from tests.common import parking_map, random_poses

nmaps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 128
jobs = []
ctx0 = hpp.Context(0)
for k in range(nmaps):
    pm = parking_map(256, seed=k)
    cx = hpp.Context(0)
    m = hpp.Map(cx, pm, 0.2)
    cand = random_poses(6 * npairs + 2000, 256, 256, margin=3.0, seed=100 + k)
    free = cand[m.collide_circles(cand) == 0][:2 * npairs]      # the device verdicts (bit-exact with the oracle)
    assert len(free) == 2 * npairs
    jobs.append(dict(ctx=cx, map=m, starts=free[:npairs], goals=free[npairs:]))
for j in jobs:
    j["map"].plan_batch(hpp.VARIANT_RS, j["starts"], j["goals"])

def work(j):
    j["plans"], _ = j["map"].plan_batch(hpp.VARIANT_RS, j["starts"], j["goals"])
for rep in range(2):
    th = [threading.Thread(target=work, args=(j,)) for j in jobs]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    it = np.concatenate([j["plans"]["iterations"] for j in jobs])
    ok = sum(int((j["plans"]["n_states"] > 1).sum()) for j in jobs)
    print(f"{nmaps} maps x {npairs} queries: {dt*1e3:.0f} ms -> {nmaps*npairs/dt:.0f} q/s  solved={ok}/{nmaps*npairs} "
          f"iters mean={it.mean():.0f} p50={np.median(it):.0f} cap-hit={(it >= 3500).sum()}")
