"""GPU parity: K6 (successor expansion), K7 (Reeds-Shepp shot), K8 (Dubins shot) and the a7/a11 unit hooks, through the
C ABI vs the CPU oracle.  The device evaluates sin/cos/atan2/... with CUDA's libdevice (<= 2 ulp) while the reference
uses glibc, so trigonometry-dependent doubles are compared within 1e-9 absolute (north_star: path cost within 1e-4
relative); everything discrete (validity flags, verdicts, chosen candidate, sample counts, cell ids, integer-valued
costs) must be identical."""
import numpy as np
import pytest

from tests.common import block_map, random_poses

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _free_poses(om, n, seed, margin=3.0, variant=0):
    from oracle import pyorc as O
    out = []
    k = 0
    while len(out) < n:
        p = random_poses(4 * n, om.W, om.H, margin=margin, seed=seed + 1000 * k)
        v = O.collide_circles(om, p) if variant == 0 else O.collide_sat(om, p)
        out.extend(p[v == 0])
        k += 1
    return np.array(out[:n])


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("W,H,p,blk,seed", [(150, 150, 0.03, 3, 1), (256, 200, 0.02, 4, 2)])
def test_expand_matches_oracle(ctx, variant, W, H, p, blk, seed):
    import hybrid_path_planning_b200 as hpp
    from hybrid_path_planning_b200.capi import make_nodes
    from oracle import pyorc as O
    occ = block_map(W, H, p, blk, seed)
    m = hpp.Map(ctx, occ, 0.2)
    om = O.Map(occ, 0.2)
    rng = np.random.default_rng(seed)
    poses = _free_poses(om, 200, seed, variant=variant)
    parents = make_nodes(poses)
    parents["cost"] = rng.uniform(0, 500, len(poses))
    parents["steer"] = rng.choice([0.0, 0.3, -0.6, 0.1], len(poses))
    parents["direction"] = rng.choice([1, -1], len(poses))
    kids = m.expand(variant, parents)
    ncmd = m.num_commands(variant)
    assert ncmd == (26 if variant == 0 else 13)
    n_valid = 0
    for i, pr in enumerate(parents):
        want = O.expand(om, variant, (pr["x"], pr["y"], pr["heading"]), float(pr["cost"]), float(pr["steer"]),
                        int(pr["direction"]))
        assert len(want) == ncmd
        for c in range(ncmd):
            g, w = kids[i, c], want[c]
            assert g["steer"] == w["steer"] and g["direction"] == w["direction"]
            assert bool(g["valid"]) == bool(w["valid"]), (i, c)
            np.testing.assert_allclose(g["traj"][:4, :3], w["traj"][:, :3], rtol=0, atol=TOL)
            assert np.array_equal(g["traj"][:4, 3:], w["traj"][:, 3:])
            if w["valid"]:
                assert g["cost"] == w["cost"]          # PathCost has no trigonometry: bit-exact
                n_valid += 1
    assert n_valid > 100
    m.close()


def _cmp_paths(got, want):
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a["types"] == b["types"]
        np.testing.assert_allclose(a["lengths"], b["lengths"], rtol=0, atol=1e-9)
        assert abs(a["L"] - b["L"]) < 1e-9
        assert len(a["x"]) == len(b["x"])
        np.testing.assert_allclose(a["x"], b["x"], rtol=0, atol=TOL)
        np.testing.assert_allclose(a["y"], b["y"], rtol=0, atol=TOL)
        np.testing.assert_allclose(a["yaw"], b["yaw"], rtol=0, atol=TOL)
        assert np.array_equal(a["dir"], b["dir"])          # PATH::directions


def test_rs_all_paths_matches_oracle(ctx):
    from oracle import pyorc as O
    rng = np.random.default_rng(4)
    maxc = np.tan(0.6) / 2.75
    n_paths = 0
    for k in range(400):
        s = (rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-np.pi, np.pi))
        g = (rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-np.pi, np.pi))
        if k % 5 == 0:
            g = (s[0] + rng.uniform(-3, 3), s[1] + rng.uniform(-3, 3), g[2])
        want = O.rs_all_paths(s, g, maxc, 1.0)
        got = ctx.rs_all_paths(s, g, maxc, 1.0)
        _cmp_paths(got, want)
        n_paths += len(want)
    assert n_paths > 2000


def test_rs_all_paths_other_step(ctx):
    from oracle import pyorc as O
    rng = np.random.default_rng(5)
    for k in range(40):
        s = (rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-np.pi, np.pi))
        g = (rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-np.pi, np.pi))
        _cmp_paths(ctx.rs_all_paths(s, g, 0.5, 0.2), O.rs_all_paths(s, g, 0.5, 0.2))   # STEP_SIZE default 0.2


def test_dubins_path_matches_oracle(ctx):
    from oracle import pyorc as O
    rng = np.random.default_rng(6)
    curv = np.tan(0.6) / 2.75
    for k in range(400):
        s = (rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-np.pi, np.pi))
        g = (rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-np.pi, np.pi))
        if k % 5 == 0:
            g = (s[0] + rng.uniform(-2, 2), s[1] + rng.uniform(-2, 2), g[2])
        want = O.dubins_path(s, g, curv)
        got = ctx.dubins_path(s, g, curv)
        assert got["mode"] == want["mode"]
        assert abs(got["length"] - want["length"]) < 1e-9
        assert len(got["x"]) == len(want["x"])
        np.testing.assert_allclose(got["x"], want["x"], rtol=0, atol=TOL)
        np.testing.assert_allclose(got["y"], want["y"], rtol=0, atol=TOL)
        np.testing.assert_allclose(got["yaw"], want["yaw"], rtol=0, atol=TOL)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("W,H,p,blk,seed", [(150, 150, 0.03, 3, 1), (256, 256, 0.02, 4, 2), (300, 300, 0.005, 3, 3)])
def test_shots_match_oracle(ctx, variant, W, H, p, blk, seed):
    import hybrid_path_planning_b200 as hpp
    from hybrid_path_planning_b200.capi import make_nodes
    from oracle import pyorc as O
    occ = block_map(W, H, p, blk, seed)
    m = hpp.Map(ctx, occ, 0.2)
    om = O.Map(occ, 0.2)
    n = 300
    starts = _free_poses(om, n, seed, variant=variant)
    goals = _free_poses(om, n, seed + 77, variant=variant)
    nodes = make_nodes(starts, cost=12.5)
    fn = m.rs_shot if variant == 0 else m.dubins_shot
    ofn = O.rs_shot if variant == 0 else O.dubins_shot
    shots, samples = fn(nodes, goals)
    found = 0
    for i in range(n):
        w = ofn(om, tuple(starts[i]), 12.5, tuple(goals[i]))
        g = shots[i]
        assert bool(g["found"]) == bool(w["found"]), i
        if variant == 0:
            assert g["n_paths"] == w["n_paths"]
        if w["found"]:
            found += 1
            assert g["chosen"] == w["chosen"]
            assert abs(g["cost"] - w["cost"]) <= 1e-9 * max(1.0, abs(w["cost"]))
            assert g["n_samples"] == len(w["traj"])
            off = int(g["sample_offset"])
            tr = samples[off:off + int(g["n_samples"])]
            np.testing.assert_allclose(tr[:, :3], w["traj"][:, :3], rtol=0, atol=TOL)
            assert np.array_equal(tr[:, 3:], w["traj"][:, 3:])
    assert found >= 1
    # one shared goal (goal_stride 0) gives the same answers as the per-start goal array
    shots1, _ = fn(nodes, goals[:1])
    shotsN, _ = fn(nodes, np.repeat(goals[:1], n, 0))
    assert np.array_equal(shots1["found"], shotsN["found"]) and np.array_equal(shots1["cost"], shotsN["cost"])
    m.close()


def test_shot_capacity_error(ctx):
    import hybrid_path_planning_b200 as hpp
    from hybrid_path_planning_b200.capi import make_nodes
    m = hpp.Map(ctx, np.zeros((200, 200), np.int8), 0.2)
    starts = np.array([[-5.0, -5.0, 0.0]])
    goals = np.array([[6.0, 4.0, 1.0]])
    shots, samples = m.rs_shot(make_nodes(starts), goals)
    assert shots[0]["found"] and shots[0]["n_samples"] > 4
    with pytest.raises(hpp.HppError):
        m.rs_shot(make_nodes(starts), goals, samples_cap=2)
    m.close()
