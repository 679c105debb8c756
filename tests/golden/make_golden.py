"""Generates tests/golden/*.npz.  Run in the build container only (needs /root/reference for the PNG maps and the
compiled reference in oracle/_ref, plus cv2 for the OpenCV-defined pieces):

    python tests/golden/make_golden.py

Sources of truth:
  * edt_*, inflate_*      cv2 4.13 with IPP disabled (cv::distanceTransform / threshold / dilate are what the reference
                          calls at RS/src/Grid_map.cpp:64 and MI/src/map_inflate_node.cpp:127-135)
  * ref_*                 the reference's own translation units compiled by oracle/build.sh (oracle/_ref)
  * maps                  the reference's PNG fixtures converted with grid_map_publisher's convention
Nothing here is reference source code; only small numeric arrays are committed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

from oracle import pyorc as O  # noqa: E402
from tests.common import block_map  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HPP_REF", "/root/reference")
cv2.ipp.setUseIPP(False)
rng = np.random.default_rng(20261019)


def png_to_occ(name):
    """grid_map_publisher (GP/src/grid_publisher.cpp:32-61,78-86): image rows map to grid_map x, the "obstacle" layer is
    255 where the PNG is black; toOccupancyGrid writes cells in reversed order: pixel (r,c) -> cell x = rows-1-r,
    y = cols-1-c, value 100 for black pixels, 0 otherwise.  OccupancyGrid width = image rows."""
    img = cv2.imread(os.path.join(REF, "grid_map_publisher", name), cv2.IMREAD_GRAYSCALE)
    rows, cols = img.shape
    occ_xy = np.where(img[::-1, ::-1] < 128, 100, 0).astype(np.int8)   # [x][y]
    return np.ascontiguousarray(occ_xy.T)                               # [y][x], W = rows, H = cols


def cv_layers(occ, res):
    free = ~((occ > 20) | (occ < 0))
    d = cv2.distanceTransform(np.ascontiguousarray(free.T.astype(np.uint8)), cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
    d = cv2.flip(d, -1)
    return (d * np.float32(res)).astype(np.float32).T.reshape(-1)


def cv_inflate(occ, res, radius):
    o = occ.astype(np.int32).copy()
    o[o < 0] = 100
    gray = (255 - np.round(o * 255.0 / 100.0)).astype(np.uint8)[::-1]
    r = max(1, int(round(radius / res)))
    _, b = cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
    el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * r + 1, 2 * r + 1), (r, r))
    out = cv2.bitwise_not(cv2.dilate(cv2.bitwise_not(b), el))
    return (100 - np.round(out[::-1].astype(np.float64) * 100.0 / 255.0)).astype(np.int8)


def main():
    res = float(np.float32(0.2))
    maps = {n.split(".")[0]: png_to_occ(n) for n in ("map_parking.png", "map_maze.png", "map_empty.png", "gridmap.png")}
    np.savez_compressed(os.path.join(HERE, "maps.npz"), **maps)

    # ---- layers / inflate (cv2)
    g = {}
    cases = {"parking": maps["map_parking"], "maze": maps["map_maze"], "blk": block_map(96, 70, 0.1, 3, 1),
             "mixed": rng.choice(np.array([-1, 0, 10, 21, 50, 100], np.int8), size=(64, 80), p=[.05, .75, .05, .05, .05, .05]),
             "empty": np.zeros((20, 30), np.int8)}
    for k, occ in cases.items():
        g[f"occ_{k}"] = occ
        g[f"dist_{k}"] = cv_layers(occ, 0.2)
        for rad in (0.2, 0.6):
            g[f"infl_{k}_{int(rad * 10)}"] = cv_inflate(occ, res, rad)
    np.savez_compressed(os.path.join(HERE, "layers_cv2.npz"), **g)

    # ---- compiled reference
    g = {}
    occ = maps["map_maze"]
    infl = cv_inflate(occ, res, 0.2)
    blk = block_map(120, 90, 0.04, 3, 2)
    g["occ_maze_infl"], g["occ_blk"] = infl, blk
    for name, o, goal in (("maze", infl, (25, 25)), ("blk", blk, (60, 45)), ("parking", maps["map_parking"], (15, 30))):
        o = o.copy()
        o[goal[1], goal[0]] = 0
        g[f"heur_occ_{name}"] = o
        g[f"heur_goal_{name}"] = np.array(goal)
        g[f"heur_rs_{name}"] = O.Ref(0, o).heuristic(goal)
        g[f"heur_db_{name}"] = O.Ref(1, o).heuristic(goal)
    maxc = np.tan(0.6) / 2.75
    pairs = np.stack([rng.uniform(-20, 20, 24), rng.uniform(-20, 20, 24), rng.uniform(-np.pi, np.pi, 24),
                      rng.uniform(-20, 20, 24), rng.uniform(-20, 20, 24), rng.uniform(-np.pi, np.pi, 24)], 1)
    pairs[::4, 3:5] = pairs[::4, 0:2] + rng.uniform(-3, 3, (6, 2))
    g["pairs"] = pairs
    for i, p in enumerate(pairs):
        ps = O.ref_rs_all_paths(p[:3], p[3:], maxc, 1.0)
        g[f"rs_{i}_n"] = np.array(len(ps))
        for k, q in enumerate(ps):
            g[f"rs_{i}_{k}_types"] = np.frombuffer(q["types"].encode(), np.uint8)
            g[f"rs_{i}_{k}_lengths"] = q["lengths"]
            g[f"rs_{i}_{k}_xyyaw"] = np.stack([q["x"], q["y"], q["yaw"]], 1)
        d = O.ref_dubins_path(p[:3], p[3:], maxc)
        g[f"db_{i}_mode"] = np.frombuffer(d["mode"].encode(), np.uint8)
        g[f"db_{i}_length"] = np.array(d["length"])
        g[f"db_{i}_xyyaw"] = np.stack([d["x"], d["y"], d["yaw"]], 1)
    # collision / expansion / shots / runs on the block map
    r0, r1 = O.Ref(0, blk), O.Ref(1, blk)
    ext_x, ext_y = 120 * res / 2, 90 * res / 2
    poses = np.stack([rng.uniform(-ext_x - 1, ext_x + 1, 400), rng.uniform(-ext_y - 1, ext_y + 1, 400),
                      rng.uniform(-np.pi, np.pi, 400)], 1)
    g["poses"] = poses
    g["sat_verdicts"] = r1.collide(poses)
    om = O.Map(blk)
    free = poses[(O.collide_circles(om, poses) == 0) & (g["sat_verdicts"] == 0)]
    g["free"] = free
    for v, r in ((0, r0), (1, r1)):
        for i in range(4):
            kids = r.expand(tuple(free[i]), 10.0 * i, 0.1 * i, 1 if i % 2 == 0 else -1)
            g[f"exp{v}_{i}_traj"] = np.array([k["traj"] for k in kids]).reshape(len(kids), 20)
            g[f"exp{v}_{i}_meta"] = np.array([[k["cost"], k["steer"], k["direction"]] for k in kids]).reshape(len(kids), 3)
        for i in range(12):
            s = r.shot(tuple(free[i]), 5.0, tuple(free[20 + i]))
            g[f"shot{v}_{i}_found"] = np.array(s["found"])
            g[f"shot{v}_{i}_cost"] = np.array(s["cost"])
            g[f"shot{v}_{i}_traj"] = s["traj"]
        for i in range(8):
            g[f"run{v}_{i}_traj"] = r.run(tuple(free[i]), tuple(free[40 + i]))
    np.savez_compressed(os.path.join(HERE, "reference.npz"), **g)
    for f in ("maps.npz", "layers_cv2.npz", "reference.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
