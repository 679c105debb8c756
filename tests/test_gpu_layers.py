"""GPU parity: K1 (obstacle + distance layers) and K2 (map_inflate) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from tests.common import block_map, bits_equal, cfg3_map

pytestmark = pytest.mark.gpu


def _layers_case(ctx, occ, res=0.2):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    m = hpp.Map(ctx, occ, res)
    obst, dist = m.layers()
    o_obst, o_dist = O.edt_layers(occ, np.float32(res))
    assert bits_equal(obst, o_obst), "obstacle layer differs"
    bad = np.flatnonzero(dist.view(np.uint32) != o_dist.view(np.uint32))
    assert bad.size == 0, f"{bad.size} distance cells differ, first {bad[:5]}: {dist[bad[:5]]} vs {o_dist[bad[:5]]}"
    m.close()


@pytest.mark.parametrize("W,H,p,blk,seed", [
    (30, 60, 0.2, 2, 1), (50, 50, 0.1, 1, 2), (64, 64, 0.3, 8, 3), (701, 710, 0.05, 3, 4), (257, 129, 0.3, 8, 5),
    (1024, 512, 0.3, 8, 6), (333, 1000, 0.01, 1, 7), (100, 7, 0.2, 1, 8), (7, 100, 0.2, 1, 9), (1, 1, 0.0, 1, 10),
])
def test_layers_bit_exact(ctx, W, H, p, blk, seed):
    _layers_case(ctx, block_map(W, H, p, blk, seed))


def test_layers_unknown_and_threshold_values(ctx):
    rng = np.random.default_rng(3)
    occ = rng.choice(np.array([-1, 0, 10, 20, 21, 50, 100], np.int8), size=(120, 200), p=[.05, .7, .05, .05, .05, .05, .05])
    _layers_case(ctx, occ)


def test_layers_empty_and_full_maps(ctx):
    _layers_case(ctx, np.zeros((40, 56), np.int8))          # no obstacle at all: cv2's constant
    _layers_case(ctx, np.full((40, 56), 100, np.int8))      # all obstacle: zeros
    occ = np.zeros((300, 260), np.int8)
    occ[150, 7] = 100                                        # one obstacle: distances beyond the shared-memory halo
    _layers_case(ctx, occ)


def test_layers_other_resolutions(ctx):
    occ = block_map(200, 100, 0.1, 2, 11)
    for res in (1.0, 0.05, 0.3):
        _layers_case(ctx, occ, res)


def test_layers_update_reuses_map(ctx):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    a, b = block_map(128, 96, 0.2, 4, 1), block_map(128, 96, 0.1, 2, 2)
    m = hpp.Map(ctx, a, 0.2)
    m.update(b)
    _, dist = m.layers()
    assert bits_equal(dist, O.edt_layers(b, np.float32(0.2))[1])
    m.close()


def test_layers_cfg3_size_properties(ctx):
    """4096^2 (BASELINE cfg3): checked against the oracle on a 512-row band's worth of exact squared distances, plus
    size-independent properties on the whole layer: zero exactly on obstacle cells, 180-degree index reversal,
    and 1-Lipschitz in cells between 4-neighbours."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ, _ = cfg3_map(4096, seed=1)
    m = hpp.Map(ctx, occ, 0.2)
    obst, dist = m.layers()
    n = occ.size
    res_f = np.float32(0.2)
    edt = dist[::-1].reshape(4096, 4096)            # undo distance[k] = edt[n-1-k]
    blocked = (occ > 20) | (occ < 0)
    assert np.array_equal(edt == 0, blocked)
    assert bits_equal(obst.reshape(4096, 4096), (~blocked).astype(np.float32))
    d = edt / res_f
    assert np.all(np.abs(np.diff(d, axis=0)) <= 1.0 + 1e-4) and np.all(np.abs(np.diff(d, axis=1)) <= 1.0 + 1e-4)
    # exact check of a band through the oracle's integer EDT (rows 1024..1535 with a 256-row apron)
    lo, hi, ap = 1024, 1536, 256
    d2 = O.edt_sq_exact(occ[lo - ap:hi + ap])[ap:-ap]
    assert d2.max() < ap * ap, "apron too small for an exact band check"
    want = (np.sqrt(d2.astype(np.float32)) * res_f).astype(np.float32)
    assert bits_equal(edt[lo:hi], want)
    m.close()


@pytest.mark.parametrize("W,H,p,blk,seed,radius", [
    (50, 50, 0.1, 1, 1, 0.2), (50, 50, 0.1, 1, 1, 0.4), (120, 77, 0.05, 2, 2, 1.0), (64, 64, 0.02, 1, 3, 2.6),
    (40, 40, 0.0, 1, 4, 0.2), (300, 500, 0.03, 3, 5, 0.9), (30, 60, 1.0, 1, 6, 0.2),
])
def test_inflate_bit_exact(ctx, W, H, p, blk, seed, radius):
    from oracle import pyorc as O
    occ = block_map(W, H, p, blk, seed)
    res = float(np.float32(0.2))
    got = ctx.inflate(occ, res, radius)
    assert bits_equal(got, O.inflate(occ, res, radius))


def test_inflate_multivalued_otsu(ctx):
    from oracle import pyorc as O
    rng = np.random.default_rng(5)
    res = float(np.float32(0.2))
    for vals in ([0, 30, 60, 100, -1], [0, 100, -1], [0, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100]):
        occ = rng.choice(np.array(vals, np.int8), size=(90, 130))
        assert bits_equal(ctx.inflate(occ, res, 0.6), O.inflate(occ, res, 0.6))


@pytest.mark.parametrize("rows,cols,seed", [(16, 16, 1), (37, 91, 2), (300, 257, 3), (1, 5, 4)])
def test_image_to_occupancy_bit_exact(ctx, rows, cols, seed):
    """f1: device image -> OccupancyGrid conversion == the oracle's restatement of grid_map_publisher's, all gray levels"""
    from oracle import pyorc as O
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
    if rows * cols >= 256:
        img.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)
    got = ctx.image_to_occupancy(img)
    assert got.shape == (cols, rows)
    assert bits_equal(got, O.image_to_occ(img))


def test_layers_sparse_maps_far_path(ctx):
    """Distances far beyond the column kernel's 32-row halo: the flagged tiles are finished by the block-minimum far
    kernel.  A nearly empty map (hundreds of cells to the nearest obstacle), a map that is dense in one corner and
    empty elsewhere (flagged and unflagged tiles side by side), a single obstacle cell in a corner, and rows without any
    obstacle; bit-exact against the oracle (= cv2, tests/test_cpu_oracle_golden.py) and, for the compact layer K4 reads,
    through the collision verdicts of poses all over the map."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    from tests.common import random_poses
    rng = np.random.default_rng(3)
    a = np.zeros((1100, 1300), np.int8)
    for _ in range(7):
        y, x = rng.integers(0, 1095), rng.integers(0, 1295)
        a[y:y + 3, x:x + 4] = 100
    b = np.zeros((900, 1000), np.int8)
    b[:200, :256] = block_map(256, 200, 0.3, 8, 5)
    c = np.zeros((700, 333), np.int8)
    c[699, 0] = 100
    d = np.zeros((300, 500), np.int8)
    d[100, :] = 100
    d[250, 40] = -1
    for occ in (a, b, c, d):
        H, W = occ.shape
        m = hpp.Map(ctx, occ, 0.2)
        om = O.Map(occ, 0.2)
        obst, dist = m.layers()
        assert bits_equal(obst, om.obstacle)
        assert bits_equal(dist, om.distance), f"{(dist != om.distance).sum()} cells differ ({W}x{H})"
        poses = random_poses(20000, W, H, margin=0.5, seed=1)
        assert np.array_equal(m.collide_circles(poses), O.collide_circles(om, poses))
        m.close()
