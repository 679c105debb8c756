"""GPU: the C++ host shells (hybrid_path_planning_b200/host: Grid_map, HybridAstar, CarData, InflateOccupancy) driven
like the reference's ROS 2 node drives its classes, compared with the oracle.  The demo binary is built by
__graft_entry__.build() (hybrid_path_planning_b200/host/build.sh)."""
import os
import subprocess

import numpy as np
import pytest

from tests.common import block_map, random_poses

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "hybrid_path_planning_b200", "shell_demo")


def _run_demo(tmp_path, occ, variant, start, goal):
    raw = tmp_path / "occ.raw"
    occ.tofile(raw)
    H, W = occ.shape
    args = [DEMO, str(raw), str(W), str(H), str(variant)] + [repr(float(v)) for v in (*start, *goal)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout.splitlines()


@pytest.mark.parametrize("variant", [0, 1])
def test_host_shells_against_oracle(tmp_path, variant):
    from oracle import pyorc as O
    assert os.path.exists(DEMO), "shell_demo missing: run __graft_entry__.build()"
    occ = block_map(120, 90, 0.04, 3, seed=2)
    om = O.Map(occ, 0.2)
    poses = random_poses(400, 120, 90, margin=1.5, seed=8)
    free = poses[(O.collide_circles(om, poses) == 0) & (O.collide_sat(om, poses) == 0)]
    start, goal = tuple(free[0]), tuple(free[5])
    out = _run_demo(tmp_path, occ, variant, start, goal)
    i = 0
    assert out[i] == "collide 0 0"
    i += 1
    assert out[i] == f"inflated {int((O.inflate(occ, float(np.float32(0.2)), 0.2) > 0).sum())}"
    i += 1
    want = O.run(om, variant, start, goal, quirks=False)
    n = int(out[i].split()[1])
    assert n == len(want["traj"])
    got = np.array([[float(v) for v in ln.split()] for ln in out[i + 1:i + 1 + n]])
    i += 1 + n
    assert np.array_equal(got[:, 3:], want["traj"][:, 3:])                  # cell indices bit-exact
    assert np.allclose(got[:, :3], want["traj"][:, :3], rtol=0, atol=1e-9)  # device libm vs glibc: <= a few ulp
    # heuristic through the shell == oracle fixed point
    cell = (int((goal[0] - om.ox) / om.res), int((goal[1] - om.oy) / om.res))
    h = O.heuristic_rs(occ, cell, (0, 0), quirks=False)[0] if variant == 0 else O.heuristic_dubins(occ, cell)
    fin = h < 1e300
    _, size, finite, total = out[i].split()
    assert int(size) == h.size and int(finite) == int(fin.sum())
    assert float(total) == pytest.approx(float(h[fin].sum()), rel=1e-12)
    i += 1
    kids = [k for k in O.expand(om, variant, start, 0.0, 0.0, 1) if k["valid"]]
    assert out[i] == f"children {len(kids)}"
    rows = np.array([[float(v) for v in ln.split()] for ln in out[i + 1:i + 1 + len(kids)]])
    i += 1 + len(kids)
    for r, k in zip(rows, kids):
        assert np.allclose(r[:3], k["traj"][-1][:3], rtol=0, atol=1e-12)
        assert r[3] == pytest.approx(k["cost"], rel=1e-12) and r[4] == k["steer"] and int(r[5]) == k["direction"]
    shot = (O.rs_shot if variant == 0 else O.dubins_shot)(om, start, 0.0, goal)
    f = out[i].split()
    assert int(f[1]) == int(shot["found"])
    if shot["found"]:
        assert float(f[2]) == pytest.approx(shot["cost"], rel=1e-9) and int(f[3]) == len(shot["traj"])
    i += 1
    # motionCommands / PathCost: the reference's own GetnextNeighbours cost for every command (colliding ones included:
    # the oracle's expansion carries the cost of every command)
    every = O.expand(om, variant, start, 0.0, 0.0, 1)
    steer, direction = O.motion_commands(variant)
    assert out[i] == f"commands {len(steer)}"
    rows = np.array([[float(v) for v in ln.split()] for ln in out[i + 1:i + 1 + len(steer)]])
    i += 1 + len(steer)
    assert np.array_equal(rows[:, 0], steer) and np.array_equal(rows[:, 1], direction)
    for r, k in zip(rows, every):
        # RS/src/HybridAstar.cpp:80-107 for the start node (Cost_path 0, steering 0, direction 1): simulationLength 4,
        # reverse 10, directionChange 250 (RS) / 200 (Dubins), steerAngle 15, steerAngleChange 20
        want_cost = (4.0 if r[1] == 1 else 4.0 * 10.0) + (0.0 if r[1] == 1 else (250.0 if variant == 0 else 200.0)) + \
            abs(r[0]) * 15.0 + abs(r[0] - 0.0) * 20.0
        assert r[2] == pytest.approx(want_cost, rel=1e-12)
        if k["valid"]:                       # the oracle (like the reference) prices only collision-free rollouts
            assert r[2] == pytest.approx(k["cost"], rel=1e-12)
    # reeds_shepp_path / dubins_path class shells
    maxc = np.tan(0.6) / 2.75
    want_all = O.rs_all_paths(start, goal, maxc, 1.0)
    assert out[i] == f"rs_paths {len(want_all)}"
    for ln, w in zip(out[i + 1:i + 1 + len(want_all)], want_all):
        f = ln.split()
        assert f[0] == w["types"] and float(f[1]) == pytest.approx(w["L"], abs=1e-9) and int(f[2]) == len(w["x"])
        assert float(f[3]) == pytest.approx(w["x"].sum(), abs=1e-7) and float(f[4]) == pytest.approx(w["y"].sum(), abs=1e-7)
        assert int(f[5]) == int(w["dir"].sum())
    i += 1 + len(want_all)
    best = min(reversed(want_all), key=lambda w: w["L"])       # `<=` in the reference: the last among equals
    f = out[i].split()
    assert f[0] == "rs_optimal" and float(f[1]) == pytest.approx(best["L"], abs=1e-9) and int(f[2]) == len(best["x"])
    i += 1
    d = O.dubins_path(start, goal, maxc, 0.1)
    f = out[i].split()
    assert f[0] == "dubins" and f[1] == d["mode"] and float(f[2]) == pytest.approx(d["length"], abs=1e-9)
    assert int(f[3]) == len(d["x"]) and float(f[4]) == pytest.approx(-1.0 + 2 * np.pi, abs=1e-15)
