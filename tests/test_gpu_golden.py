"""GPU parity against the COMMITTED golden fixtures (tests/golden/*.npz: cv2 with IPP off + the reference's own
translation units compiled in the build container; generator: tests/golden/make_golden.py).  No oracle call here: the
CUDA path through the C ABI is compared with stored reference outputs directly.  Bit-exact for layers, inflation,
verdicts, heuristic fields, cell indices and trig-free costs; <= 1e-9 absolute where the device's libm differs from
glibc by an ulp or two (sampled poses); path cost within 1e-4 relative."""
import os

import numpy as np
import pytest

from tests.common import bits_equal

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9


@pytest.fixture(scope="module")
def layers():
    return np.load(os.path.join(G, "layers_cv2.npz"))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(G, "reference.npz"))


@pytest.mark.parametrize("case", ["parking", "maze", "blk", "mixed", "empty"])
def test_layers_and_inflation_vs_cv2_golden(ctx, layers, case):
    import hybrid_path_planning_b200 as hpp
    occ = layers[f"occ_{case}"]
    m = hpp.Map(ctx, occ, 0.2)
    obstacle, distance = m.layers()
    if case != "empty":                      # cv2's value for an image without zero pixels is not pinned (A.5)
        assert bits_equal(distance, layers[f"dist_{case}"])
    assert bits_equal(obstacle, np.where((occ > 20) | (occ < 0), 0.0, 1.0).astype(np.float32).reshape(-1))
    for rad in (0.2, 0.6):
        assert bits_equal(ctx.inflate(occ, float(np.float32(0.2)), rad), layers[f"infl_{case}_{int(rad * 10)}"])
    m.close()


@pytest.mark.parametrize("name", ["maze", "blk", "parking"])
def test_heuristic_fields_vs_reference_golden(ctx, ref, name):
    import hybrid_path_planning_b200 as hpp
    occ = ref[f"heur_occ_{name}"]
    goal = tuple(int(v) for v in ref[f"heur_goal_{name}"])
    m = hpp.Map(ctx, occ, 0.2)
    rs, rep = m.heuristic_rs(goal, (0, 0))
    assert rep.h1_progress_break == 0 and rep.h2_bucket_clamp == 0
    assert bits_equal(rs, ref[f"heur_rs_{name}"].reshape(rs.shape))
    db, _ = m.heuristic_dubins(goal)
    assert bits_equal(db, ref[f"heur_db_{name}"].reshape(db.shape))
    m.close()


def test_sat_verdicts_vs_reference_golden(ctx, ref):
    import hybrid_path_planning_b200 as hpp
    m = hpp.Map(ctx, ref["occ_blk"], 0.2)
    assert np.array_equal(m.collide_sat(ref["poses"]), ref["sat_verdicts"])
    m.close()


def test_analytic_paths_vs_reference_golden(ctx, ref):
    maxc = np.tan(0.6) / 2.75
    for i, p in enumerate(ref["pairs"]):
        got = ctx.rs_all_paths(p[:3], p[3:], maxc, 1.0)
        assert len(got) == int(ref[f"rs_{i}_n"]), i
        for k, q in enumerate(got):
            assert q["types"].encode() == ref[f"rs_{i}_{k}_types"].tobytes()
            np.testing.assert_allclose(q["lengths"], ref[f"rs_{i}_{k}_lengths"], rtol=0, atol=TOL)
            want = ref[f"rs_{i}_{k}_xyyaw"]
            assert len(q["x"]) == len(want)
            np.testing.assert_allclose(np.stack([q["x"], q["y"], q["yaw"]], 1), want, rtol=0, atol=TOL)
        d = ctx.dubins_path(p[:3], p[3:], maxc)
        assert d["mode"].encode() == ref[f"db_{i}_mode"].tobytes().rstrip(b"\0")
        assert d["length"] == pytest.approx(float(ref[f"db_{i}_length"]), abs=TOL)
        want = ref[f"db_{i}_xyyaw"]
        assert len(d["x"]) == len(want)
        np.testing.assert_allclose(np.stack([d["x"], d["y"], d["yaw"]], 1), want, rtol=0, atol=TOL)


@pytest.mark.parametrize("variant", [0, 1])
def test_expansion_shots_runs_vs_reference_golden(ctx, ref, variant):
    import hybrid_path_planning_b200 as hpp
    from hybrid_path_planning_b200.capi import make_nodes
    m = hpp.Map(ctx, ref["occ_blk"], 0.2)
    free = ref["free"]
    for i in range(4):
        parents = make_nodes(free[i:i + 1], cost=10.0 * i, steer=0.1 * i, direction=1 if i % 2 == 0 else -1)
        kids = [k for k in m.expand(variant, parents)[0] if k["valid"]]
        traj, meta = ref[f"exp{variant}_{i}_traj"], ref[f"exp{variant}_{i}_meta"]
        assert len(kids) == len(traj)
        for k, t, mt in zip(kids, traj, meta):
            t = t.reshape(4, 5)
            np.testing.assert_allclose(k["traj"][:4, :3], t[:, :3], rtol=0, atol=TOL)
            assert np.array_equal(k["traj"][:4, 3:], t[:, 3:])
            assert k["cost"] == mt[0] and k["steer"] == mt[1] and k["direction"] == int(mt[2])
    shot = m.rs_shot if variant == 0 else m.dubins_shot
    starts = make_nodes(free[:12], cost=5.0)
    shots, samples = shot(starts, free[20:32])
    for i in range(12):
        assert int(shots[i]["found"]) == int(ref[f"shot{variant}_{i}_found"])
        if shots[i]["found"]:
            want = ref[f"shot{variant}_{i}_traj"]
            assert abs(shots[i]["cost"] - float(ref[f"shot{variant}_{i}_cost"])) <= 1e-4 * max(1.0, float(ref[f"shot{variant}_{i}_cost"]))
            off, n = int(shots[i]["sample_offset"]), int(shots[i]["n_samples"])
            assert n == len(want)
            np.testing.assert_allclose(samples[off:off + n, :3], want[:, :3], rtol=0, atol=TOL)
            assert np.array_equal(samples[off:off + n, 3:], want[:, 3:])
    plans, traj = m.plan_batch(variant, free[:8], free[40:48])
    for i in range(8):
        want = ref[f"run{variant}_{i}_traj"]
        assert (plans[i]["n_states"] > 1) == (len(want) > 1)
        if plans[i]["n_states"] == len(want):
            off = int(plans[i]["traj_offset"])
            np.testing.assert_allclose(traj[off:off + len(want), :3], want[:, :3], rtol=0, atol=TOL)
            assert np.array_equal(traj[off:off + len(want), 3:], want[:, 3:])
    m.close()
