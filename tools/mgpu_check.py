"""Multi-GPU check of the C ABI's multi-GPU layer (csrc/mgpu.cu) on real GPUs: run under torchrun, one rank per GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_check.py
The NCCL communicator lives inside libhpp_b200.so (hpp_comm_create); torch.distributed only ships the 128-byte id.
Checks, against the single-GPU result computed on every rank:
  * hpp_collide_circles_sharded   (pose batch split, verdicts all-gathered)
  * hpp_plan_batch_sharded        (queries on several maps split, plan headers all-gathered; also with fewer queries
                                   than ranks)
  * hpp_distance_layer_sharded    (row bands, halo exchange of the pass-1 rows g over ncclSend/Recv): a dense map (first
                                   halo enough), a sparse map with an odd row count (halo has to be doubled, bands of
                                   unequal height), and a nearly empty map (falls back to gathering all rows)
Prints one JSON line on rank 0; exit code 0 iff every rank agrees everywhere."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybrid_path_planning_b200 as hpp  # noqa: E402
from tests.common import block_map, cfg3_map, parking_map, random_poses  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = hpp.Context(local)
idt = torch.zeros(hpp.capi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(hpp.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
comm = hpp.Comm(ctx, bytes(idt.cpu().numpy().tobytes()), rank, world)
out = {"world": world}
ok = True

# ---- pose batch (cfg4 shape, smaller)
side = 2048
occ, _ = cfg3_map(side, seed=2)
m = hpp.Map(ctx, occ, 0.2)
poses = random_poses(2_000_001, side, side, margin=2.0, seed=3)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
verd = comm.collide_circles_sharded(m, poses)
t1 = time.perf_counter()
out["collide_equal"] = bool(np.array_equal(verd, m.collide_circles(poses)))
out["collide_sharded_s"] = t1 - t0
ok &= out["collide_equal"]

# ---- queries on three maps (cfg5 shape), interleaved
occs = [parking_map(256, seed=k) for k in range(3)]
maps = [hpp.Map(ctx, o, 0.2) for o in occs]
starts, goals, moq = [], [], []
for k, mm in enumerate(maps):
    cand = random_poses(3000, 256, 256, margin=3.0, seed=7 + k)
    free = cand[mm.collide_circles(cand) == 0][:86]
    starts.append(free[:43]); goals.append(free[43:86]); moq += [k] * 43
starts, goals, moq = np.vstack(starts), np.vstack(goals), np.array(moq, np.uint32)
perm = np.random.default_rng(1).permutation(len(moq))
starts, goals, moq = starts[perm], goals[perm], moq[perm]
plans, traj, (lo, hi) = comm.plan_batch_sharded(maps, moq, hpp.VARIANT_RS, starts, goals)
ref_plans, ref_traj = hpp.plan_batch_multi(maps, moq, hpp.VARIANT_RS, starts, goals)
same = all(np.array_equal(plans[f], ref_plans[f]) for f in ("n_states", "iterations", "via_shot", "status", "cost"))
for q in range(lo, hi):      # this rank's trajectories
    a = traj[int(plans[q]["traj_offset"]):int(plans[q]["traj_offset"]) + int(plans[q]["n_states"])]
    b = ref_traj[int(ref_plans[q]["traj_offset"]):int(ref_plans[q]["traj_offset"]) + int(ref_plans[q]["n_states"])]
    same &= bool(np.array_equal(a, b))
out["plans_equal"] = bool(same)
p1, _, span = comm.plan_batch_sharded(maps, moq[:1], hpp.VARIANT_RS, starts[:1], goals[:1])   # fewer queries than ranks
out["one_query_equal"] = bool(p1[0]["cost"] == ref_plans[0]["cost"] and p1[0]["n_states"] == ref_plans[0]["n_states"])
ok &= out["plans_equal"] and out["one_query_equal"]

# ---- distance layer in row bands
cases = {"dense": (cfg3_map(2048, seed=4)[0], 64), "sparse_odd": (block_map(1536, 2049, 0.002, 4, seed=11), 8),
         "nearly_empty": (np.where(np.arange(1000 * 1200).reshape(1200, 1000) == 777, 100, 0).astype(np.int8), 16)}
for name, (big, halo) in cases.items():
    H, W = big.shape
    lo, hi = hpp.shard_range(H, world, rank)
    band_occ = torch.from_numpy(big[lo:hi].copy()).to(dev)
    band = torch.empty((hi - lo) * W, dtype=torch.float32, device=dev)
    full = torch.empty(H * W, dtype=torch.float32, device=dev)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    rep = comm.distance_layer_sharded(band_occ, W, H, 0.2, band, full, halo=halo)
    t1 = time.perf_counter()
    ref = hpp.Map(ctx, big, 0.2)
    _, want = ref.layers()                                   # single-GPU layer, grid_map layout
    ref.close()
    eq_full = bool(full.cpu().numpy().tobytes() == want.tobytes())
    eq_band = bool(band.cpu().numpy().tobytes() == want[::-1].reshape(H, W)[lo:hi].tobytes())
    out[f"edt_{name}"] = {"equal_full": eq_full, "equal_band": eq_band, "rounds": rep.rounds, "halo": rep.halo_rows,
                          "gathered_all": rep.gathered_all, "s": t1 - t0,
                          "max_cells": float(want.max() / np.float32(0.2))}
    ok &= eq_full and eq_band

flags = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
out["all_ranks_ok"] = bool(flags.item() == 1)
if rank == 0:
    print(json.dumps(out), flush=True)
for mm in maps:
    mm.close()
m.close(); comm.close(); ctx.close()
dist.destroy_process_group()
sys.exit(0 if out["all_ranks_ok"] else 1)
