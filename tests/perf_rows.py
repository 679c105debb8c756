"""Throughput of the remaining hot-path entry points (K2 inflate, K5 SAT footprint, K6 expansion, K7/K8 shots) on the
device, one line per kernel, next to a bounded single-core CPU sample of the oracle port.  Not the bench (bench.py
measures the headline); this fills the per-row numbers of DESIGN.md / profiles.  usage: python tests/perf_rows.py"""
import os, sys, time, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import hybrid_path_planning_b200 as hpp
from hybrid_path_planning_b200.capi import make_nodes
from oracle import pyorc as O
from tests.common import cfg3_map, block_map, parking_map, random_poses

ctx = hpp.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
out = {}


def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); ctx.sync(); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts[1:])


# K2 inflate, 4096^2, r = 3 px (0.6 m at 0.2 m)
side = 4096
occ, _ = cfg3_map(side, seed=1)
d_occ = torch.from_numpy(occ).cuda(); d_out = torch.empty_like(d_occ)
t = timed(lambda: ctx.inflate_dev(d_occ, side, side, float(np.float32(0.2)), 0.6, d_out))
small = occ[:1024, :1024].copy()
t0 = time.perf_counter(); O.inflate(small, float(np.float32(0.2)), 0.6); tc = time.perf_counter() - t0
out["K2_inflate"] = {"ms": t * 1e3, "mcells_per_s": side * side / t / 1e6, "gb_per_s_algorithmic_2B": 2 * side * side / t / 1e9,
                     "cpu_port_mcells_per_s_1core": small.size / tc / 1e6}

# K5 SAT footprint: 10^6 poses on a 2048^2 map with 5 % blocks
occ2 = block_map(2048, 2048, 0.05, 8, seed=5)
m = hpp.Map(ctx, occ2, 0.2)
poses = random_poses(1_000_000, 2048, 2048, margin=6.0, seed=1)
d_p = torch.from_numpy(poses).cuda(); d_v = torch.empty(len(poses), dtype=torch.uint8, device="cuda")
t = timed(lambda: m.collide_sat_dev(d_p, len(poses), d_v))
om = O.Map(occ2[:512, :512], 0.2)
ps = random_poses(300, 512, 512, margin=6.0, seed=2)
t0 = time.perf_counter(); O.collide_sat(om, ps); tc = time.perf_counter() - t0
out["K5_collide_sat"] = {"ms": t * 1e3, "mposes_per_s": len(poses) / t / 1e6, "colliding": float(d_v.float().mean().item()),
                         "cpu_port_mposes_per_s_1core": len(ps) / tc / 1e6}

# K6 expansion: 10^5 parents (RS: 26 children x 4 footprint checks each; Dubins: 13 x 4 SAT checks)
pm = parking_map(256, seed=0)
mp = hpp.Map(ctx, pm, 0.2)
cand = random_poses(400_000, 256, 256, margin=3.0, seed=3)
free = cand[mp.collide_circles(cand) == 0][:100_000]
parents = make_nodes(free)
for variant, name in ((0, "K6_expand_rs"), (1, "K6_expand_dubins")):
    t = timed(lambda: mp.expand(variant, parents), reps=3)          # host buffers: includes H2D/D2H of the children
    omp = O.Map(pm, 0.2)
    t0 = time.perf_counter()
    for i in range(200):
        O.expand(omp, variant, tuple(free[i]), 0.0, 0.0, 1)
    tc = time.perf_counter() - t0
    from hybrid_path_planning_b200.capi import CHILD_DTYPE
    nc = mp.num_commands(variant)
    d_par = torch.from_numpy(parents.view(np.uint8).reshape(len(parents), -1).copy()).cuda()
    d_kid = torch.empty(len(parents) * nc * CHILD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    td = timed(lambda: mp.expand_dev(variant, d_par, len(parents), d_kid))
    out[name] = {"ms_device": td * 1e3, "mparents_per_s_device": len(parents) / td / 1e6,
                 "mchecks_per_s_device": len(parents) * nc * 4 / td / 1e6, "child_bytes": CHILD_DTYPE.itemsize,
                 "ms_e2e": t * 1e3, "mparents_per_s_e2e": len(parents) / t / 1e6,
                 "cpu_port_mparents_per_s_1core": 200 / tc / 1e6}

# K7 / K8 analytic shots: 20 000 start nodes towards one goal each
starts = make_nodes(free[:20_000], cost=1.0)
goals = free[20_000:40_000]
for fn, cpu, name in ((mp.rs_shot, O.rs_shot, "K7_rs_shot"), (mp.dubins_shot, O.dubins_shot, "K8_dubins_shot")):
    t = timed(lambda: fn(starts, goals, samples_cap=1 << 23), reps=3)
    omp = O.Map(pm, 0.2)
    t0 = time.perf_counter()
    for i in range(200):
        cpu(omp, tuple(free[i]), 1.0, tuple(goals[i]))
    tc = time.perf_counter() - t0
    out[name] = {"ms_e2e": t * 1e3, "kshots_per_s_e2e": len(starts) / t / 1e3, "cpu_port_kshots_per_s_1core": 200 / tc / 1e3}
print(json.dumps(out, indent=1))
