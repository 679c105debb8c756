"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

What shards: independent units only -- pose batches (cfg4), start/goal queries and their goal fields (cfg5), and row
bands of a large map for the distance layer (cfg3/cfg4 maps).  Units are split contiguously over the ranks; there is no
data-path collective during the compute.  Collectives are used only where the path has a real exchange step:
  * result gather (verdicts / plan headers):   one all_gather of the per-rank slices, padded to the longest slice;
  * distance-layer row bands:                  one halo exchange of occupancy rows with the two neighbouring ranks
                                               (send/recv), widened and repeated only if a band reports a distance
                                               that reaches past its halo.
A single heuristic field does not shard (the wavefront crosses any band boundary thousands of times): replicas only,
shard by goal instead.

The compute callables are injected so the same host logic runs on the GPU (NCCL, the C-ABI library) and in the CPU test
suite (gloo, world_size 2, the oracle as the compute stand-in -- tests only)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous split of n units: rank r owns [lo, hi), sizes differ by at most one (first n % world ranks larger)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(n: int, world: int) -> list[int]:
    return [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_slices(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all ranks contribute their contiguous slice (rows) and receive the concatenation [n_total, ...]"""
    world, rank = _world(group)
    if world == 1:
        return local
    counts = shard_counts(n_total, world)
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    longest = max(counts)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def sharded_map(units: np.ndarray, compute, device="cpu", group=None) -> np.ndarray:
    """split `units` (rows) over the ranks, run compute(rows) -> array with one row per unit on each slice, gather.
    compute is the C-ABI call on the GPU ranks (e.g. Map.collide_circles) -- every rank returns the full result."""
    world, rank = _world(group)
    lo, hi = shard_range(len(units), world, rank)
    out = np.asarray(compute(units[lo:hi]))
    if out.shape[0] != hi - lo:
        raise ValueError(f"compute returned {out.shape[0]} rows for {hi - lo} units")
    t = torch.from_numpy(np.ascontiguousarray(out)).to(device)
    return gather_slices(t, len(units), group).cpu().numpy()


def sharded_queries(starts: np.ndarray, goals: np.ndarray, plan, device="cpu", group=None):
    """cfg5: queries q in [g*Q/G, (g+1)*Q/G) on rank g; plan(starts, goals) -> (structured plans, traj);
    plan headers are gathered (viewed as bytes), trajectories stay on the rank that produced them."""
    world, rank = _world(group)
    lo, hi = shard_range(len(starts), world, rank)
    plans, traj = plan(starts[lo:hi], goals[lo:hi])
    # explicit item size: a rank whose slice is empty (fewer queries than ranks) still takes part in the gather
    raw = torch.from_numpy(np.ascontiguousarray(plans).view(np.uint8).reshape(len(plans), plans.dtype.itemsize).copy()).to(device)
    allp = gather_slices(raw, len(starts), group).cpu().numpy()
    return allp.reshape(-1).view(plans.dtype), (lo, hi, traj)


# ------------------------------------------------------------------------------------------- distance-layer row bands
def band_with_halo(occ_band: torch.Tensor, halo: int, group=None, min_rows: int | None = None) -> tuple[torch.Tensor, int, int]:
    """occ_band: this rank's rows [rows, W] int8 (device tensor for NCCL, CPU tensor for gloo).  Exchanges `halo` rows
    with the ranks above and below and returns (band + halo rows, rows received from below-index neighbour, rows
    received from above-index neighbour).  Rank r holds rows after rank r-1 (y grows with the rank).
    The halo is clamped with min_rows = total_rows // world, the SAME number on every rank (bands of a shard_range
    split differ by one row when total_rows % world != 0: a clamp with the local height would make neighbours post
    sends and receives of different sizes)."""
    world, rank = _world(group)
    rows = occ_band.shape[0]
    h = min(halo, rows if min_rows is None else min_rows)
    if world == 1:
        return occ_band, 0, 0
    ops, lo_buf, hi_buf = [], None, None
    if rank > 0:
        lo_buf = torch.empty((h,) + tuple(occ_band.shape[1:]), dtype=occ_band.dtype, device=occ_band.device)
        ops.append(dist.P2POp(dist.isend, occ_band[:h].contiguous(), rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, lo_buf, rank - 1, group))
    if rank < world - 1:
        hi_buf = torch.empty((h,) + tuple(occ_band.shape[1:]), dtype=occ_band.dtype, device=occ_band.device)
        ops.append(dist.P2POp(dist.isend, occ_band[rows - h:].contiguous(), rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, hi_buf, rank + 1, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    parts = ([lo_buf] if lo_buf is not None else []) + [occ_band] + ([hi_buf] if hi_buf is not None else [])
    return torch.cat(parts, 0), (h if lo_buf is not None else 0), (h if hi_buf is not None else 0)


def sharded_distance_band(occ_band: torch.Tensor, total_rows: int, resolution: float, edt_cells, halo: int = 64,
                          group=None):
    """Exact Euclidean distance for this rank's band of a map of total_rows rows split with shard_range.
    edt_cells(occ[rows, W]) -> float32 [rows, W]: distances in CELLS of a stand-alone map (sqrtf of the exact integer
    squared distance -- the K1 kernel on the GPU ranks).  A band result is exact when no distance in it exceeds the
    distance to the first row beyond its halo (an obstacle out there cannot be nearer); otherwise the halo is doubled
    and the exchange repeated (the ranks agree through one all_reduce of the flag); once the halo would pass the
    neighbouring band the occupancy is all-gathered instead.  Returns metres (float32), band rows only, not flipped."""
    world, rank = _world(group)
    rows = occ_band.shape[0]
    res = float(np.float32(resolution))
    min_rows = total_rows // max(world, 1)
    while True:
        ext, below, above = band_with_halo(occ_band, halo, group, min_rows)
        band = edt_cells(ext)[below:below + rows]
        idx = torch.arange(rows, device=band.device, dtype=torch.float32)
        limit = torch.full((rows,), float("inf"), device=band.device)
        if rank > 0:
            limit = torch.minimum(limit, idx + (below + 1))
        if rank < world - 1:
            limit = torch.minimum(limit, (rows - 1 - idx) + (above + 1))
        unresolved = (band > limit[:, None]).any().to(torch.int32).reshape(1)
        if world > 1:
            dist.all_reduce(unresolved, op=dist.ReduceOp.MAX, group=group)
        if int(unresolved.item()) == 0:
            return band * res
        if halo >= min_rows:
            break
        halo *= 2
    full = gather_slices(occ_band, total_rows, group)
    lo, hi = shard_range(total_rows, world, rank)
    return edt_cells(full)[lo:hi] * res
