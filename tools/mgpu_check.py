"""Multi-GPU check of hybrid_path_planning_b200/shard.py on real GPUs (NCCL): run under torchrun, one rank per GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_check.py
Checks (against the single-GPU result computed on every rank): sharded circle-collision verdicts with all_gather,
sharded queries with the plan-header gather, and the distance layer of a map split into row bands with the halo
exchange (send/recv) -- K1 runs on each band + halo through the C ABI.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp  # noqa: E402
from hybrid_path_planning_b200 import shard  # noqa: E402
from tests.common import block_map, cfg3_map, random_poses  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = hpp.Context(local)
out = {"world": world}

# ---- pose batch (cfg4 shape, smaller): every rank holds the map, poses are split, verdicts gathered
side = 2048
occ, _ = cfg3_map(side, seed=2)
m = hpp.Map(ctx, occ, 0.2)
poses = random_poses(2_000_001, side, side, margin=2.0, seed=3)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
verd = shard.sharded_map(poses, lambda p: m.collide_circles(p), device=dev)
torch.cuda.synchronize(); t1 = time.perf_counter()
full = m.collide_circles(poses)
out["collide_equal"] = bool(np.array_equal(verd, full))
out["collide_sharded_s"] = t1 - t0

# ---- queries (cfg5 shape): split, planned per rank, plan headers gathered
from tests.common import parking_map  # noqa: E402
pm = parking_map(256, seed=1)
mq = hpp.Map(ctx, pm, 0.2)
cand = random_poses(3000, 256, 256, margin=3.0, seed=7)
free = cand[mq.collide_circles(cand) == 0][:130]
starts, goals = free[:65], free[65:130]
allp, (lo, hi, _) = shard.sharded_queries(starts, goals, lambda s, g: mq.plan_batch(hpp.VARIANT_RS, s, g), device=dev)
ref_plans, _ = mq.plan_batch(hpp.VARIANT_RS, starts, goals)
out["plans_equal"] = bool(np.array_equal(allp["n_states"], ref_plans["n_states"]) and
                          np.array_equal(allp["cost"], ref_plans["cost"]))

# ---- distance layer in row bands with halo exchange (NCCL send/recv of device tensors)
big = block_map(1536, 2048, 0.002, 4, seed=11)          # sparse obstacles: distances of tens of cells
H, W = big.shape


def edt_cells(occ_t):
    """K1 on a stand-alone map at resolution 1 -> distances in cells, un-flipped [rows, W] (device tensor)"""
    rows = occ_t.shape[0]
    mm = hpp.Map(ctx, (W, rows), 1.0, device_occ=occ_t.contiguous())
    _, dptr = mm.layers_dev()
    ctx.sync()
    d = torch.empty(rows * W, dtype=torch.float32, device=dev)
    torch.cuda.current_stream().synchronize()
    import ctypes
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(dptr), ctypes.c_size_t(rows * W * 4), 3)
    mm.close()
    return torch.flip(d, [0]).reshape(rows, W)


lo, hi = shard.shard_range(H, world, rank)
band = torch.from_numpy(big[lo:hi].copy()).to(dev)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
d_band = shard.sharded_distance_band(band, H, 0.2, edt_cells, halo=8)
torch.cuda.synchronize(); t1 = time.perf_counter()
d_full = edt_cells(torch.from_numpy(big).to(dev)) * float(np.float32(0.2))
out["edt_band_equal"] = bool(torch.equal(d_band, d_full[lo:hi]))
out["edt_max_cells"] = float(d_full.max().item() / float(np.float32(0.2)))
out["edt_band_s"] = t1 - t0
flags = torch.tensor([int(out["collide_equal"]), int(out["plans_equal"]), int(out["edt_band_equal"])], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
out["all_ranks_ok"] = bool(flags.min().item() == 1)
if rank == 0:
    print(json.dumps(out), flush=True)
m.close(); mq.close(); ctx.close()
dist.destroy_process_group()
sys.exit(0 if out["all_ranks_ok"] else 1)
