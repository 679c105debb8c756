"""CPU (no GPU): pins the C oracle (oracle/liborc.so) against the committed golden vectors in tests/golden/:
  * layers_cv2.npz  -- cv2 4.13 (IPP off): cv::distanceTransform / threshold / dilate, the library calls the reference
                       makes at RS/src/Grid_map.cpp:64 and MI/src/map_inflate_node.cpp:127-135
  * reference.npz   -- outputs of the reference's own translation units compiled by oracle/build.sh (oracle/_ref):
                       heuristic fields, calc_all_paths, calc_dubins_path, SAT verdicts, expansions, shots, whole runs
The generating script is tests/golden/make_golden.py.  Everything is compared bit-for-bit except trajectories of the
libm-heavy analytic paths, which the oracle computes with the same glibc and therefore also match exactly here."""
import os

import numpy as np
import pytest

from oracle import pyorc as O
from tests.common import bits_equal

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def layers():
    return np.load(os.path.join(G, "layers_cv2.npz"))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(G, "reference.npz"))


@pytest.fixture(scope="module", autouse=True)
def _built():
    O.build()


@pytest.mark.parametrize("case", ["parking", "maze", "blk", "mixed"])
def test_distance_layer_matches_cv2(layers, case):
    occ = layers[f"occ_{case}"]
    obstacle, distance = O.edt_layers(occ, np.float32(0.2))
    assert bits_equal(distance, layers[f"dist_{case}"])
    # obstacle layer: RS/src/Grid_map.cpp:35-51 (1 = free, 0 = data > 20 or unknown), layout x + y*W
    want = np.where((occ > 20) | (occ < 0), 0.0, 1.0).astype(np.float32).reshape(-1)
    assert bits_equal(obstacle, want)


@pytest.mark.parametrize("case", ["parking", "maze", "blk", "mixed", "empty"])
@pytest.mark.parametrize("rad", [0.2, 0.6])
def test_inflate_matches_cv2(layers, case, rad):
    occ = layers[f"occ_{case}"]
    got = O.inflate(occ, float(np.float32(0.2)), rad)
    assert bits_equal(got, layers[f"infl_{case}_{int(rad * 10)}"])


@pytest.mark.parametrize("name", ["maze", "blk", "parking"])
def test_heuristics_match_compiled_reference(ref, name):
    occ = ref[f"heur_occ_{name}"]
    goal = tuple(int(v) for v in ref[f"heur_goal_{name}"])
    rs, q = O.heuristic_rs(occ, goal, (0, 0), quirks=True)
    assert bits_equal(rs, ref[f"heur_rs_{name}"].reshape(rs.shape))
    db = O.heuristic_dubins(occ, goal)
    assert bits_equal(db, ref[f"heur_db_{name}"].reshape(db.shape))
    # the order-independent fixed point (what the device computes) equals the reference here: no quirk fires
    assert not q.h1_fired and not q.h2_clamped
    assert bits_equal(O.heuristic_rs(occ, goal, (0, 0), quirks=False)[0], rs)


def test_rs_all_paths_match_compiled_reference(ref):
    maxc = np.tan(0.6) / 2.75
    for i, p in enumerate(ref["pairs"]):
        got = O.rs_all_paths(p[:3], p[3:], maxc, 1.0)
        assert len(got) == int(ref[f"rs_{i}_n"])
        for k, q in enumerate(got):
            assert q["types"].encode() == ref[f"rs_{i}_{k}_types"].tobytes()
            assert bits_equal(q["lengths"], ref[f"rs_{i}_{k}_lengths"])
            assert bits_equal(np.stack([q["x"], q["y"], q["yaw"]], 1), ref[f"rs_{i}_{k}_xyyaw"])


def test_dubins_path_matches_compiled_reference(ref):
    maxc = np.tan(0.6) / 2.75
    for i, p in enumerate(ref["pairs"]):
        d = O.dubins_path(p[:3], p[3:], maxc)
        assert d["mode"].encode() == ref[f"db_{i}_mode"].tobytes().rstrip(b"\0")
        assert d["length"] == float(ref[f"db_{i}_length"])
        assert bits_equal(np.stack([d["x"], d["y"], d["yaw"]], 1), ref[f"db_{i}_xyyaw"])


def test_sat_verdicts_match_compiled_reference(ref):
    om = O.Map(ref["occ_blk"])
    assert np.array_equal(O.collide_sat(om, ref["poses"]), ref["sat_verdicts"])


@pytest.mark.parametrize("variant", [0, 1])
def test_expand_shots_runs_match_compiled_reference(ref, variant):
    om = O.Map(ref["occ_blk"])
    free = ref["free"]
    for i in range(4):
        kids = O.expand(om, variant, tuple(free[i]), 10.0 * i, 0.1 * i, 1 if i % 2 == 0 else -1)
        kids = [k for k in kids if k["valid"]]           # the reference only returns the collision-free children
        traj = np.array([k["traj"] for k in kids]).reshape(len(kids), 20)
        meta = np.array([[k["cost"], k["steer"], k["direction"]] for k in kids]).reshape(len(kids), 3)
        assert bits_equal(traj, ref[f"exp{variant}_{i}_traj"])
        assert bits_equal(meta, ref[f"exp{variant}_{i}_meta"])
    shot = O.rs_shot if variant == 0 else O.dubins_shot
    for i in range(12):
        s = shot(om, tuple(free[i]), 5.0, tuple(free[20 + i]))
        assert int(s["found"]) == int(ref[f"shot{variant}_{i}_found"])
        if s["found"]:
            assert s["cost"] == float(ref[f"shot{variant}_{i}_cost"])
            assert bits_equal(s["traj"], ref[f"shot{variant}_{i}_traj"].reshape(s["traj"].shape))
    for i in range(8):
        r = O.run(om, variant, tuple(free[i]), tuple(free[40 + i]), quirks=True)
        want = ref[f"run{variant}_{i}_traj"]
        assert bits_equal(r["traj"], want.reshape(r["traj"].shape))


def test_sanity_constants():
    """constants probed from the reference build"""
    steer, d = O.motion_commands(0)
    assert len(steer) == 26 and steer[0] == 0.6 and steer[12] == 8.3266726846886740532e-17
    assert steer[24] == -0.59999999999999986677 and list(d[:2]) == [1.0, -1.0]
    steer1, _ = O.motion_commands(1)
    assert len(steer1) == 13
    assert np.tan(0.6) / 2.75 == 0.24877702121516085
    assert float(np.float32(0.2)) == 0.20000000298023224


def test_exact_edt_property():
    """distance layer == sqrtf(exact integer squared distance) * float(res), reversed"""
    rng = np.random.default_rng(3)
    occ = (rng.random((57, 83)) < 0.05).astype(np.int8) * 100
    d2 = O.edt_sq_exact(occ)
    _, dist = O.edt_layers(occ, np.float32(0.2))
    want = (np.sqrt(d2.astype(np.float32)) * np.float32(0.2)).astype(np.float32).reshape(-1)[::-1]
    assert bits_equal(dist, np.ascontiguousarray(want))


def test_image_to_occupancy_convention():
    """f1 (GP/src/grid_publisher.cpp:32-50,78-86): the oracle's image -> OccupancyGrid conversion reproduces the committed
    fixture maps from their (binary) images -- width = image rows, both axes reversed, black = 100 -- and maps gray
    levels through grid_map's float arithmetic with truncation."""
    maps = np.load(os.path.join(G, "maps.npz"))
    for k in maps.files:
        occ = maps[k]
        img = np.where(occ.T[::-1, ::-1] > 0, 0, 255).astype(np.uint8)      # pixel (i, j) <- cell (rows-1-i, cols-1-j)
        got = O.image_to_occ(img)
        assert got.shape == occ.shape and bits_equal(got, occ)
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16)
    got = O.image_to_occ(ramp)
    want = np.empty(256, np.int8)
    for v in range(256):
        mv = np.float32(255.0) + np.float32(-255.0) * (np.float32(v) / np.float32(255.0))
        val = min(max(np.float32(0.0), mv / np.float32(255.0)), np.float32(1.0)) * np.float32(100.0)
        i, j = divmod(v, 16)
        want[255 - (j * 16 + i)] = np.int8(int(val))
    assert bits_equal(got.reshape(-1), want)
    assert got.reshape(-1)[255] == 100 and got.reshape(-1)[0] == 0


def test_edt_layers_wide_maps_against_live_cv2():
    """Maps wider than 4096 cells: a float32 evaluation of the parabola intersections (q*q + f[q] beyond 24 bits) puts
    cells on the wrong parabola; cv2 4.13 (IPP off) returns the exact transform there and so must the oracle (the
    8192^2 cfg4 map differs in 11 % of its cells otherwise).  cv2 is part of the image, here and on the GPU box."""
    cv2 = pytest.importorskip("cv2")
    from tests.common import cfg3_map
    rng = np.random.default_rng(3)
    sparse = np.zeros((40, 9000), np.int8)
    for _ in range(25):
        y, x = rng.integers(0, 38), rng.integers(0, 8990)
        sparse[y:y + 2, x:x + 3] = 100
    tall = np.ascontiguousarray(sparse.T)
    strip = cfg3_map(8192, seed=2)[0][4000:4096]          # 96 rows of the cfg4 map, full width
    for occ in (sparse, tall, strip):
        ob, di = O.edt_layers(occ, np.float32(0.2))
        ob2, di2 = O.cv2_layers(occ, 0.2)
        assert bits_equal(ob, ob2) and bits_equal(di, di2), f"{(di != di2).sum()} cells differ"
