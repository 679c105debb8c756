"""GPU parity on BASELINE.json configs[0] and configs[1] -- the reference's own fixture maps (tests/golden/maps.npz, the
PNGs of grid_map_publisher converted with its convention):
  cfg1  reeds_shepp_hybrid_astar on map_parking.png at 0.2 m, default vehicle, single start/goal queries
  cfg2  dubins_hybrid_star on map_maze.png at 0.2 m with a 0.2 m inflation radius
plus the two variants SURVEY 8(d) asks for because almost nothing is solvable at 0.2 m on those tiny maps: the same maps
at the packages' default 1.0 m resolution, and gridmap.png at 0.2 m (multi-iteration searches).
Outcome-level parity per query against the REFERENCE-EQUIVALENT run -- the oracle with the reference's sequential
heuristic quirks ON (bit-identical to the compiled reference, tests/test_cpu_oracle_vs_ref.py), and the compiled
reference itself (oracle/_ref) where it is present: same success, path cost within 1e-4 relative, collision-free path,
cell indices of identical replays bit-exact.  On gridmap.png (497 710 cells > 70 000) the reference's progress-check
quirk H1 (RS/src/HybridAstar.cpp:452-472) fires in normal use: the test counts the heuristic cells on which the
reference's field differs from the exact fixed point the device returns, and asserts that hpp_quirk_report flags it."""
import os

import numpy as np
import pytest

from tests.common import random_poses

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pairs(om, variant, n, seed, margin):
    from oracle import pyorc as O
    out = []
    k = 0
    while len(out) < 2 * n and k < 50:
        p = random_poses(400, om.W, om.H, res=om.res, margin=margin, seed=seed + 97 * k)
        v = O.collide_circles(om, p) if variant == 0 else O.collide_sat(om, p)
        out.extend(p[v == 0])
        k += 1
    out = np.array(out[:2 * n])
    return out[:n], out[n:]


def _check(ctx, occ, res, variant, starts, goals):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    m = hpp.Map(ctx, occ, res)
    om = O.Map(occ, res)
    plans, traj = m.plan_batch(variant, starts, goals)
    solved = identical = 0
    ref = O.Ref(variant, occ, res) if O.ref_available() else None
    for i in range(len(starts)):
        w = O.run(om, variant, tuple(starts[i]), tuple(goals[i]), quirks=True)   # = the reference, quirks and all
        if ref is not None and i < 8:
            rt = ref.run(tuple(starts[i]), tuple(goals[i]))                      # the compiled reference itself
            assert rt.shape == w["traj"].shape and np.array_equal(rt, w["traj"]), i
        g = plans[i]
        assert g["status"] == 0
        assert (g["n_states"] > 1) == (len(w["traj"]) > 1), i
        assert abs(g["cost"] - w["cost"]) <= 1e-4 * max(1.0, abs(w["cost"])), i
        off = int(g["traj_offset"])
        tr = traj[off:off + int(g["n_states"])]
        if g["n_states"] > 1:
            solved += 1
            hit = O.collide_circles(om, tr[:, :3]) if variant == 0 else O.collide_sat(om, tr[:, :3])
            assert not hit.any()
        if g["iterations"] == w["iterations"] and g["n_states"] == len(w["traj"]):
            identical += 1
            np.testing.assert_allclose(tr[:, :3], w["traj"][:, :3], rtol=0, atol=1e-9)
            assert np.array_equal(tr[:, 3:], w["traj"][:, 3:])
    m.close()
    return solved, identical


def test_cfg1_rs_on_map_parking(ctx):
    maps = np.load(os.path.join(G, "maps.npz"))
    occ = maps["map_parking"]
    from oracle import pyorc as O
    # 0.2 m: the map is 6 m x 12 m; includes the solvable pair found by the survey's oracle scan
    om = O.Map(occ, 0.2)
    s, g = _pairs(om, 0, 40, seed=1, margin=0.3)
    s = np.vstack([s, [[-0.7, -4.9, np.pi]]])
    g = np.vstack([g, [[-0.1, -4.7, np.pi]]])
    solved, identical = _check(ctx, occ, 0.2, 0, s, g)
    assert solved >= 1 and identical >= 0.9 * len(s)
    # 1.0 m (package default, GP/config/params.yaml): multi-iteration searches
    om1 = O.Map(occ, 1.0)
    s, g = _pairs(om1, 0, 24, seed=2, margin=3.0)
    solved, identical = _check(ctx, occ, 1.0, 0, s, g)
    assert solved >= 3 and identical >= 0.85 * len(s)


def test_cfg1_rs_on_gridmap(ctx):
    maps = np.load(os.path.join(G, "maps.npz"))
    occ = maps["gridmap"]
    from oracle import pyorc as O
    om = O.Map(occ, 0.2)
    s, g = _pairs(om, 0, 6, seed=3, margin=10.0)
    s = np.vstack([s, [[66.8, -36.7, 0.785]]])           # the multi-iteration pair of SURVEY 8(d)
    g = np.vstack([g, [[21.4, -45.7, 2.356]]])
    solved, identical = _check(ctx, occ, 0.2, 0, s, g)
    assert solved >= 5 and identical >= 5
    # quirk H1 on this map: cells where the reference's field (quirks on) differs from the exact fixed point
    import hybrid_path_planning_b200 as hpp
    m = hpp.Map(ctx, occ, 0.2)
    fired = 0
    for i in range(len(s)):
        st, gl = O._state(om, *s[i]), O._state(om, *g[i])
        goal, start = (int(gl.gridx), int(gl.gridy)), (int(st.gridx), int(st.gridy))
        hq, q = O.heuristic_rs(occ, goal, start, quirks=True)
        he, _ = O.heuristic_rs(occ, goal, start, quirks=False)
        got, rep = m.heuristic_rs(goal, start)
        assert got.tobytes() == he.tobytes()                         # the device returns the exact fixed point
        differ = int((hq != he).sum())
        fin = np.isfinite(hq) & np.isfinite(he) & (he > 0)
        rel = float(np.max(np.abs(hq[fin] - he[fin]) / he[fin])) if fin.any() else 0.0
        print(f"gridmap query {i}: H1 fired={q.h1_fired}, {differ} cells differ (max rel {rel:.2e}), "
              f"hpp_quirk_report.h1_progress_break={rep.h1_progress_break}")
        if q.h1_fired:
            fired += 1
            assert differ > 0 and rep.h1_progress_break != 0        # flagged by the report whenever it fires
        else:
            assert differ == 0 or q.h2_clamped
    assert fired >= 5                                               # the quirk does fire in normal use on this map
    m.close()


def test_cfg2_dubins_on_inflated_maze(ctx):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    maps = np.load(os.path.join(G, "maps.npz"))
    occ = maps["map_maze"]
    res = float(np.float32(0.2))
    infl = ctx.inflate(occ, res, 0.2)
    assert np.array_equal(infl, O.inflate(occ, res, 0.2)) and int((infl > 0).sum()) == 777
    om = O.Map(infl, 0.2)
    s, g = _pairs(om, 1, 30, seed=4, margin=0.5)
    _, identical = _check(ctx, infl, 0.2, 1, s, g)
    assert identical >= 0.9 * len(s)
    res1 = 1.0
    infl1 = ctx.inflate(occ, res1, 0.2)
    om1 = O.Map(infl1, 1.0)
    s, g = _pairs(om1, 1, 20, seed=5, margin=4.0)
    solved, identical = _check(ctx, infl1, 1.0, 1, s, g)
    assert identical >= 0.8 * len(s)
