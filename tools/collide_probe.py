"""Timing probe for K4 on the cfg4 workload (8192^2 layer, 10^7 poses).  usage: python tools/collide_probe.py [reps] [p]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp
from tests.common import cfg3_map, random_poses
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
p_occ = float(sys.argv[2]) if len(sys.argv) > 2 else 0.30
side, npose = 8192, 10_000_000
occ, _ = cfg3_map(side, seed=2, p=p_occ)
ctx = hpp.Context(0)
m = hpp.Map(ctx, (side, side), 0.2, device_occ=torch.from_numpy(occ).cuda())
poses = torch.from_numpy(random_poses(npose, side, side, margin=2.0, seed=3)).cuda()
verd = torch.empty(npose, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
ts = []
for k in range(reps + 2):
    with torch.cuda.stream(stream):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); m.collide_circles_dev(poses, npose, verd); b.record(stream)
    ctx.sync(); ts.append(a.elapsed_time(b))
t = min(ts[2:])
import hashlib
print("verdict sha1", hashlib.sha1(verd.cpu().numpy().tobytes()).hexdigest()[:16])
print(f"K4 p={p_occ}: {t:.3f} ms  {npose / t / 1e6:.2f} Gchecks/s  {89 * npose / t / 1e6:.0f} GB/s algorithmic  colliding={verd.float().mean().item():.3f}")
