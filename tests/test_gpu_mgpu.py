"""GPU: the multi-map planner entry (hpp_plan_batch_multi) and the multi-GPU layer of the C ABI (mgpu.cu) on ONE GPU.

hpp_plan_batch_multi must return exactly what one hpp_plan_batch call per map returns.  The sharded entry points are run
on a world-size-1 NCCL communicator (NCCL works with a single rank: the all-gather / all-reduce calls are real), where
they must equal the single-GPU entry points; the band logic of the distance layer (halo clamp, unresolved flag,
widening, all-rows fall-back) is exercised for several simulated ranks by calling the same band kernels through
tools/mgpu_check.py on real multi-GPU boxes and, rank by rank, by the gloo tests of the host mirror on CPU
(tests/test_cpu_shard_gloo.py)."""
import numpy as np
import pytest

from tests.common import block_map, parking_map, random_poses

pytestmark = pytest.mark.gpu


def _free(m, n, W, H, seed):
    cand = random_poses(8 * n + 500, W, H, margin=3.0, seed=seed)
    return cand[m.collide_circles(cand) == 0][:n]


def test_plan_batch_multi_equals_per_map_calls(ctx):
    import hybrid_path_planning_b200 as hpp
    # three maps of one geometry (one launch) + one of another geometry (its own group)
    occs = [parking_map(256, seed=0), parking_map(256, seed=1), block_map(256, 256, 0.02, 4, 2), block_map(150, 130, 0.03, 3, 1)]
    maps = [hpp.Map(ctx, o, 0.2) for o in occs]
    starts, goals, moq = [], [], []
    for k, (m, o) in enumerate(zip(maps, occs)):
        H, W = o.shape
        f = _free(m, 24, W, H, seed=10 + k)
        starts.append(f[:12])
        goals.append(f[12:24])
        moq += [k] * 12
    starts, goals, moq = np.vstack(starts), np.vstack(goals), np.array(moq, np.uint32)
    perm = np.random.default_rng(3).permutation(len(moq))            # interleave the maps in the query list
    plans, traj = hpp.plan_batch_multi(maps, moq[perm], hpp.VARIANT_RS, starts[perm], goals[perm])
    assert (plans["status"] == 0).all() and (plans["n_states"] > 1).sum() >= 8
    for k, m in enumerate(maps):
        sel = np.where(moq[perm] == k)[0]
        p1, t1 = m.plan_batch(hpp.VARIANT_RS, starts[perm][sel], goals[perm][sel])
        for j, q in enumerate(sel):
            for f in ("n_states", "iterations", "via_shot", "status", "cost"):
                assert plans[q][f] == p1[j][f], (k, j, f)
            a = traj[int(plans[q]["traj_offset"]):int(plans[q]["traj_offset"]) + int(plans[q]["n_states"])]
            b = t1[int(p1[j]["traj_offset"]):int(p1[j]["traj_offset"]) + int(p1[j]["n_states"])]
            assert np.array_equal(a, b), (k, j)
    for m in maps:
        m.close()


def test_plan_batch_multi_argument_errors(ctx):
    import hybrid_path_planning_b200 as hpp
    m = hpp.Map(ctx, parking_map(256, seed=0), 0.2)
    f = _free(m, 4, 256, 256, seed=1)
    with pytest.raises(hpp.HppError):
        hpp.plan_batch_multi([m], np.array([0, 1], np.uint32), 0, f[:2], f[2:4])   # map index out of range
    plans, traj = hpp.plan_batch_multi([m], np.zeros(0, np.uint32), 0, np.zeros((0, 3)), np.zeros((0, 3)))
    assert len(plans) == 0 and len(traj) == 0
    m.close()


@pytest.fixture(scope="module")
def comm1(ctx):
    import hybrid_path_planning_b200 as hpp
    c = hpp.Comm(ctx, hpp.comm_unique_id(), 0, 1)
    yield c
    c.close()


def test_shard_range_matches_host_mirror():
    import hybrid_path_planning_b200 as hpp
    from hybrid_path_planning_b200.shard import shard_range
    for n in (0, 1, 7, 8, 201, 8192):
        for world in (1, 2, 3, 8):
            got = [hpp.shard_range(n, world, r) for r in range(world)]
            assert got == [shard_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n and all(got[i][1] == got[i + 1][0] for i in range(world - 1))


def test_sharded_entry_points_on_one_rank(ctx, comm1):
    import torch
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = block_map(300, 260, 0.05, 3, 5)
    m = hpp.Map(ctx, occ, 0.2)
    poses = random_poses(5000, 300, 260, margin=1.0, seed=2)
    assert np.array_equal(comm1.collide_circles_sharded(m, poses), m.collide_circles(poses))
    f = _free(m, 16, 300, 260, seed=4)
    plans, traj, (lo, hi) = comm1.plan_batch_sharded([m], None, hpp.VARIANT_RS, f[:8], f[8:16])
    p1, t1 = m.plan_batch(hpp.VARIANT_RS, f[:8], f[8:16])
    assert (lo, hi) == (0, 8)
    for q in range(8):   # (trajectory offsets depend on which query finishes first: compare through them)
        for fld in ("n_states", "iterations", "via_shot", "status", "cost"):
            assert plans[q][fld] == p1[q][fld], (q, fld)
        a = traj[int(plans[q]["traj_offset"]):int(plans[q]["traj_offset"]) + int(plans[q]["n_states"])]
        b = t1[int(p1[q]["traj_offset"]):int(p1[q]["traj_offset"]) + int(p1[q]["n_states"])]
        assert np.array_equal(a, b), q
    # distance layer in one band: the band result is the un-flipped layer, the gathered one the reference layout
    d_occ = torch.from_numpy(occ).cuda()
    band = torch.empty(occ.size, dtype=torch.float32, device="cuda")
    full = torch.empty(occ.size, dtype=torch.float32, device="cuda")
    rep = comm1.distance_layer_sharded(d_occ, 300, 260, 0.2, band, full)
    _, want = O.edt_layers(occ, np.float32(0.2))
    assert rep.rounds == 1 and rep.gathered_all == 0
    assert full.cpu().numpy().tobytes() == want.tobytes()
    assert band.cpu().numpy()[::-1].tobytes() == want.tobytes()
    m.close()
