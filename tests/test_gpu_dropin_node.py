"""Drop-in check end to end: the reference's UNMODIFIED planner node (its main_path_planning.cpp, compiled by
hybrid_path_planning_b200/host/build.sh from the reference checkout) running on the hpp-b200 host shells inside the stand-in
ROS of tests/ros_mock.  The harness publishes a map, /initialpose and /goal_pose to the node's callbacks and prints the
path the node publishes on /hybrid_astar_path; it must be the oracle's path for the same query, state by state.
Skipped when the binaries were not built (no reference checkout at build time)."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def _run_node(tag, occ, res, start, goal, tmp_path):
    exe = os.path.join(ROOT, "hybrid_path_planning_b200", "ref_node_" + tag)
    if not os.path.exists(exe):
        pytest.skip("ref_node_%s not built (needs the reference checkout at build time)" % tag)
    H, W = occ.shape
    res_f = np.float32(res)
    ox, oy = -W * float(res_f) / 2.0, -H * float(res_f) / 2.0
    path = os.path.join(tmp_path, "map.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<iifdd", W, H, float(res_f), ox, oy))
        f.write(np.ascontiguousarray(occ, np.int8).tobytes())
    args = [exe, path] + ["%.17g" % v for v in (*start, *goal)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    s = g = None
    states = []
    for ln in r.stdout.splitlines():
        w = ln.split()
        if w[0] == "start":
            s = tuple(float(v) for v in w[1:])
        elif w[0] == "goal":
            g = tuple(float(v) for v in w[1:])
        elif w[0] == "state":
            states.append((float(w[1]), float(w[2])))
    return s, g, np.array(states)


@pytest.mark.parametrize("tag,variant", [("rs", 0), ("db", 1)])
def test_reference_node_on_the_shells_publishes_the_oracle_path(tag, variant, tmp_path):
    from oracle import pyorc as O
    ref = np.load(os.path.join(G, "reference.npz"))
    occ, free = ref["occ_blk"], ref["free"]
    om = O.Map(occ, 0.2)
    solved = 0
    # (start, goal) indices into the golden collision-free poses; the Dubins ones were picked solvable (forward-only
    # planner: most random pairs on this map are not)
    pairs = [(i, 40 + i) for i in range(6)] if variant == 0 else [(1, 14), (12, 25), (19, 26), (0, 40), (6, 46)]
    for i, j in pairs:
        s, g, states = _run_node(tag, occ, 0.2, free[i], free[j], str(tmp_path))
        assert s is not None and g is not None
        w = O.run(om, variant, s, g, quirks=False)     # the poses exactly as the node's callbacks computed them
        assert len(states) == len(w["traj"]), (i, len(states), len(w["traj"]))
        np.testing.assert_allclose(states, w["traj"][:, :2], rtol=0, atol=1e-9)
        solved += int(len(states) > 1)
    assert solved >= 2


@pytest.mark.parametrize("name,radius", [("map_maze", 0.2), ("map_parking", 0.6), ("mixed", 0.45), ("gridmap", 0.3)])
def test_reference_map_inflate_node_on_the_cv_shim(name, radius, tmp_path):
    """MI/src/map_inflate_node.cpp, UNMODIFIED (MapCallback -> InflateMap -> CreateInflatedOccupancyGrid), with its
    cv::threshold / cv::dilate calls routed to the C ABI by host/dropin/mi/opencv2/opencv.hpp: the OccupancyGrid it
    publishes on /grid_map_inflated must equal the oracle's (= cv2's, tests/golden/layers_cv2.npz) bit for bit --
    including a map with intermediate occupancy values, where the Otsu threshold is data dependent."""
    from oracle import pyorc as O
    exe = os.path.join(ROOT, "hybrid_path_planning_b200", "ref_node_mi")
    if not os.path.exists(exe):
        pytest.skip("ref_node_mi not built (needs the reference checkout at build time)")
    if name == "mixed":
        rng = np.random.default_rng(7)
        occ = rng.choice(np.array([-1, 0, 10, 30, 60, 100], np.int8), size=(90, 130), p=[.03, .77, .05, .05, .05, .05])
    else:
        occ = np.load(os.path.join(G, "maps.npz"))[name]
    H, W = occ.shape
    res_f = np.float32(0.2)
    path, out = os.path.join(tmp_path, "map.bin"), os.path.join(tmp_path, "out.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<iifdd", W, H, float(res_f), -W * float(res_f) / 2.0, -H * float(res_f) / 2.0))
        f.write(np.ascontiguousarray(occ, np.int8).tobytes())
    r = subprocess.run([exe, path, repr(radius), out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-1500:])
    got = np.fromfile(out, np.int8).reshape(H, W)
    want = O.inflate(occ, float(res_f), radius)
    assert np.array_equal(got, want), f"{(got != want).sum()} cells differ"


def test_cv_entry_points_match_cv2(ctx):
    """hpp_cv_threshold_otsu / hpp_cv_dilate_ellipse against the OpenCV calls they replace (cv2 is part of the image)"""
    cv2 = pytest.importorskip("cv2")
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(11)
    for shape in ((97, 131), (256, 256), (33, 1000)):
        gray = rng.choice(np.array([0, 20, 77, 128, 200, 255], np.uint8), size=shape, p=[.1, .1, .2, .2, .2, .2])
        thr, b = ctx.cv_threshold_otsu(gray)
        wt, wb = cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
        assert thr == wt and np.array_equal(b, wb)
        sparse = (rng.random(shape) < 0.01).astype(np.uint8) * 255
        for rad in (0, 1, 2, 5, 9):
            el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * rad + 1, 2 * rad + 1), (rad, rad))
            want = cv2.dilate(sparse, el)
            assert np.array_equal(ctx.cv_dilate_ellipse(sparse, rad), want), (shape, rad)
            assert np.array_equal(ctx.cv_dilate_ellipse(255 - sparse, rad, complemented=True), want), (shape, rad)
