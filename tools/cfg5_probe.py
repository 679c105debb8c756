"""cfg5 on one GPU through hpp_plan_batch_multi (8 parking maps x 1024 queries), with HPP_PLAN_TIMING phase times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp  # noqa: E402
from bench import cfg5_jobs, RES  # noqa: E402

if os.environ.get("PROBE_TORCH"):  # what bench.py has around the same call: torch's CUDA state and a large allocation
    import torch
    _big = torch.empty(int(os.environ["PROBE_TORCH"]) << 30, dtype=torch.uint8, device="cuda")
    _big.fill_(1)
    torch.cuda.synchronize()
ctx = hpp.Context(0)
maps = {}


def dev_collide(k, pm, cand):
    maps[k] = hpp.Map(ctx, pm, RES)
    return maps[k].collide_circles(cand)


nmaps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
jobs = cfg5_jobs(nmaps, 1024, dev_collide)
mlist = [maps[k] for k in range(nmaps)]
starts = np.vstack([j["starts"] for j in jobs])
goals = np.vstack([j["goals"] for j in jobs])
moq = np.repeat(np.arange(nmaps, dtype=np.uint32), 1024)
traj = np.empty((1 << 21, 5))
for rep in range(4):
    t0 = time.perf_counter()
    plans, _ = hpp.plan_batch_multi(mlist, moq, hpp.VARIANT_RS, starts, goals, traj_out=traj)
    t = time.perf_counter() - t0
    print(f"pass {rep}: {1e3 * t:.1f} ms  {len(starts) / t:.0f} queries/s  solved {(plans['n_states'] > 1).sum()}  "
          f"iterations sum {plans['iterations'].sum()} max {plans['iterations'].max()}", flush=True)
