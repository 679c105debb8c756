"""BASELINE cfg5 across GPUs: 8192 independent RS Hybrid A* queries (8 synthetic parking maps x 1024 start/goal pairs),
sharded by map over the ranks (shard.shard_range), planned with hpp_plan_batch (one context per map, maps in flight
concurrently), plan headers gathered with one all_gather.  Run under torchrun (or plain python for 1 GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/cfg5_scaling.py
Prints one JSON line on rank 0: queries/s = 8192 / max-over-ranks wall time of the planning phase."""
import json, os, sys, threading, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp
from hybrid_path_planning_b200 import shard
from tests.common import parking_map, random_poses

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NMAPS, NPAIRS = 8, 1024
lo, hi = shard.shard_range(NMAPS, world, rank)
jobs = []
for k in range(lo, hi):
    pm = parking_map(256, seed=k)
    cx = hpp.Context(local)
    mm = hpp.Map(cx, pm, 0.2)
    cand = random_poses(6 * NPAIRS + 2000, 256, 256, margin=3.0, seed=100 + k)
    free = cand[mm.collide_circles(cand) == 0][:2 * NPAIRS]
    jobs.append(dict(ctx=cx, map=mm, starts=free[:NPAIRS], goals=free[NPAIRS:]))
for j in jobs:
    j["map"].plan_batch(hpp.VARIANT_RS, j["starts"], j["goals"])      # warm-up: scratch growth


def work(j):
    j["plans"], _ = j["map"].plan_batch(hpp.VARIANT_RS, j["starts"], j["goals"])


torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
th = [threading.Thread(target=work, args=(j,)) for j in jobs]
for t in th:
    t.start()
for t in th:
    t.join()
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
plans = np.concatenate([j["plans"] for j in jobs]) if jobs else np.zeros(0, hpp.capi.PLAN_DTYPE)
raw = torch.from_numpy(plans.view(np.uint8).reshape(len(plans), -1).copy()).to(dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    # every rank holds whole maps: contiguous slices of NPAIRS * maps-per-rank rows, gathered in rank order
    counts = [NPAIRS * (shard.shard_range(NMAPS, world, r)[1] - shard.shard_range(NMAPS, world, r)[0]) for r in range(world)]
    pad = torch.zeros((max(counts), raw.shape[1]), dtype=torch.uint8, device=dev)
    pad[:len(raw)] = raw
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    raw = torch.cat([p[:c] for p, c in zip(parts, counts)], 0)
allp = raw.cpu().numpy().reshape(-1).view(hpp.capi.PLAN_DTYPE)
if rank == 0:
    print(json.dumps({"workload": "cfg5", "n_gpus": world, "queries": int(len(allp)), "solved": int((allp["n_states"] > 1).sum()),
                      "seconds": float(dt.item()), "queries_per_s": len(allp) / float(dt.item()),
                      "mean_iterations": float(allp["iterations"].mean())}), flush=True)
for j in jobs:
    j["map"].close(); j["ctx"].close()
if world > 1:
    dist.destroy_process_group()
