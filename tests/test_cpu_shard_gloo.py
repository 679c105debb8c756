"""CPU (no GPU): the multi-GPU host logic of hybrid_path_planning_b200/shard.py under gloo, world_size 2.
The compute callables are the C oracle here (test infrastructure standing in for the C-ABI calls the GPU ranks make);
what is checked is the sharding, the gathers and the halo exchange: results must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hybrid_path_planning_b200 import shard
from tests.common import block_map, random_poses


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 8192, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard.shard_counts(n, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _edt_cells_oracle(occ_t):
    from oracle import pyorc as O
    occ = occ_t.cpu().numpy()
    H, W = occ.shape
    _, d = O.edt_layers(occ, np.float32(1.0))           # layout: distance[k] = edt[n-1-k]
    return torch.from_numpy(np.ascontiguousarray(d[::-1]).reshape(H, W).copy())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyorc as O
        out = {}
        # ---- pose batches (cfg4 shape): verdict gather
        occ = block_map(160, 120, 0.05, 3, seed=5)
        om = O.Map(occ, 0.2)
        poses = random_poses(1001, 160, 120, margin=1.0, seed=9)         # odd count: uneven slices
        out["verdicts"] = shard.sharded_map(poses, lambda p: O.collide_circles(om, p))
        # ---- queries (cfg5 shape): plan-header gather
        free = poses[O.collide_circles(om, poses) == 0]
        starts, goals = free[:7], free[7:14]
        dt = np.dtype([("n_states", "i4"), ("iterations", "i4"), ("cost", "f8")])

        def plan(s, g):
            plans = np.zeros(len(s), dt)
            for i in range(len(s)):
                r = O.run(om, 0, tuple(s[i]), tuple(g[i]), quirks=False)
                plans[i] = (len(r["traj"]), r["iterations"], r["cost"])
            return plans, None
        allp, (lo, hi, _) = shard.sharded_queries(starts, goals, plan)
        out["plans"] = allp
        out["span"] = (lo, hi)
        # ---- distance-layer row bands with halo exchange (halo too small on purpose: must widen, then stay exact)
        big = block_map(96, 200, 0.004, 2, seed=11)                      # sparse: distances of tens of cells
        lo, hi = shard.shard_range(big.shape[0], world, rank)
        band = torch.from_numpy(big[lo:hi].copy())
        out["band"] = shard.sharded_distance_band(band, big.shape[0], 0.2, _edt_cells_oracle, halo=4).numpy()
        out["band_span"] = (lo, hi)
        # uneven split (201 rows over 2 ranks: bands of 101 and 100 rows) with a halo that has to grow past the
        # shorter band: the clamp must be the same on both ranks or the exchange sizes disagree
        odd = block_map(96, 201, 0.0015, 2, seed=12)
        lo, hi = shard.shard_range(odd.shape[0], world, rank)
        out["band_odd"] = shard.sharded_distance_band(torch.from_numpy(odd[lo:hi].copy()), odd.shape[0], 0.2,
                                                      _edt_cells_oracle, halo=4).numpy()
        # fewer queries than ranks: the rank with the empty slice still takes part in the gather
        one, (lo1, hi1, _) = shard.sharded_queries(starts[:1], goals[:1], plan)
        out["one_plan"] = one
        out["one_span"] = (lo1, hi1)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_paths_world2():
    from oracle import pyorc as O
    O.build()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth
    occ = block_map(160, 120, 0.05, 3, seed=5)
    om = O.Map(occ, 0.2)
    poses = random_poses(1001, 160, 120, margin=1.0, seed=9)
    want = O.collide_circles(om, poses)
    for r in range(world):
        assert np.array_equal(res[r]["verdicts"], want)
    free = poses[want == 0]
    for i in range(7):
        w = O.run(om, 0, tuple(free[i]), tuple(free[7 + i]), quirks=False)
        for r in range(world):
            p = res[r]["plans"][i]
            assert p["n_states"] == len(w["traj"]) and p["iterations"] == w["iterations"] and p["cost"] == w["cost"]
    assert res[0]["span"] == (0, 4) and res[1]["span"] == (4, 7)
    big = block_map(96, 200, 0.004, 2, seed=11)
    full = _edt_cells_oracle(torch.from_numpy(big)).numpy() * float(np.float32(0.2))
    assert full.max() > 4 * 0.2 * 3, "the test map must have distances well past the initial halo"
    got = np.concatenate([res[r]["band"] for r in range(world)], 0)
    assert got.dtype == np.float32 and np.array_equal(got, full.astype(np.float32))
    odd = block_map(96, 201, 0.0015, 2, seed=12)
    full = _edt_cells_oracle(torch.from_numpy(odd)).numpy() * float(np.float32(0.2))
    assert full.max() > 0.2 * 64, "the odd-split map needs a halo past 64 rows: the next doubling (128) hits the clamp"
    got = np.concatenate([res[r]["band_odd"] for r in range(world)], 0)
    assert got.shape == (201, 96) and np.array_equal(got, full.astype(np.float32))
    assert res[0]["one_span"] == (0, 1) and res[1]["one_span"] == (1, 1)
    for r in range(world):
        assert len(res[r]["one_plan"]) == 1 and res[r]["one_plan"][0] == res[r]["plans"][0]
