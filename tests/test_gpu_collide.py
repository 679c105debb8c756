"""GPU parity: K4 (circle footprint) and K5 (SAT footprint) verdicts through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from tests.common import bits_equal, block_map, random_poses

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,p,blk,seed", [(150, 150, 0.03, 3, 1), (256, 256, 0.02, 4, 2), (301, 203, 0.1, 2, 3),
                                            (512, 512, 0.3, 8, 4)])
def test_collide_circles_bit_exact(ctx, W, H, p, blk, seed):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = block_map(W, H, p, blk, seed)
    m = hpp.Map(ctx, occ, 0.2)
    om = O.Map(occ, 0.2)
    poses = random_poses(4000, W, H, margin=-1.0, seed=seed)      # includes poses outside the map
    got = m.collide_circles(poses)
    want = O.collide_circles(om, poses)
    assert np.array_equal(got, want), f"{(got != want).sum()} verdicts differ"
    assert 0.02 < want.mean() < 0.999
    m.close()


def test_collide_circles_border_rules(ctx):
    """Poses whose circles straddle the map border exercise isInside, the linear-index aliasing of
    atPositionLinearInterpolated and the INTER_NEAREST fallback (oracle/orc_map.c:gm_at_linear)."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = np.zeros((80, 90), np.int8)
    occ[40, 45] = 100
    m = hpp.Map(ctx, occ, 0.2)
    om = O.Map(occ, 0.2)
    rng = np.random.default_rng(0)
    ext_x, ext_y = 90 * om.res / 2, 80 * om.res / 2
    side = rng.integers(0, 4, 6000)
    t = rng.uniform(-1, 1, 6000)
    d = rng.uniform(-1.2, 1.2, 6000)
    x = np.where(side == 0, -ext_x + d, np.where(side == 1, ext_x + d, t * ext_x))
    y = np.where(side == 2, -ext_y + d, np.where(side == 3, ext_y + d, t * ext_y))
    poses = np.stack([x, y, rng.uniform(-np.pi, np.pi, 6000)], 1)
    assert np.array_equal(m.collide_circles(poses), O.collide_circles(om, poses))
    m.close()


@pytest.mark.parametrize("W,H,p,blk,seed", [(150, 150, 0.03, 3, 1), (256, 256, 0.02, 4, 2), (301, 203, 0.1, 2, 3),
                                            (90, 40, 0.3, 1, 4)])
def test_collide_sat_bit_exact(ctx, W, H, p, blk, seed):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = block_map(W, H, p, blk, seed)
    m = hpp.Map(ctx, occ, 0.2)
    om = O.Map(occ, 0.2, with_layers=False)
    poses = random_poses(3000, W, H, margin=-1.0, seed=seed)
    got = m.collide_sat(poses)
    want = O.collide_sat(om, poses)
    assert np.array_equal(got, want), f"{(got != want).sum()} verdicts differ"
    m.close()


def test_collide_empty_batch_and_cells(ctx):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = block_map(64, 48, 0.2, 2, 9)
    m = hpp.Map(ctx, occ, 0.2)
    assert m.collide_circles(np.zeros((0, 3))).shape == (0,)
    assert m.collide_sat(np.zeros((0, 3))).shape == (0,)
    om = O.Map(occ, 0.2, with_layers=False)
    rng = np.random.default_rng(1)
    xy = rng.uniform(-8, 8, (5000, 2))
    cell, flags = m.cells(xy)
    cx = ((xy[:, 0] - om.ox) / om.res).astype(np.int64)   # numpy float->int truncates toward zero like static_cast<int>
    cy = ((xy[:, 1] - om.oy) / om.res).astype(np.int64)
    assert np.array_equal(cell[:, 0], cx) and np.array_equal(cell[:, 1], cy)
    inb = (cx >= 0) & (cx < 64) & (cy >= 0) & (cy < 48)
    col = ~inb
    col[inb] = occ[cy[inb], cx[inb]] > 0
    assert np.array_equal(flags & 1, inb.astype(np.uint8)) and np.array_equal((flags >> 1) & 1, col.astype(np.uint8))
    m.close()


def test_collide_circles_cfg4_size(ctx):
    """BASELINE cfg4 scale-down that the oracle finishes in seconds (2048^2 layer, 2*10^5 poses), bit-exact."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    from tests.common import cfg3_map
    occ, _ = cfg3_map(2048, seed=2)
    m = hpp.Map(ctx, occ, 0.2)
    om = O.Map(occ, 0.2)
    poses = random_poses(200000, 2048, 2048, margin=2.0, seed=5)
    got = m.collide_circles(poses)
    want = O.collide_circles(om, poses)
    assert np.array_equal(got, want)
    m.close()


def test_collide_circles_cfg4_full_size(ctx):
    """BASELINE cfg4 at its named size -- the launch bench.py times: 8192^2 distance layer (same generator and seed),
    10^7 poses.  The distance layer is bit-compared with the oracle's; a strided 1 % sample of the verdicts (every
    100th pose, 10^5 of them) is bit-compared with the oracle's isSingleStateCollisionFree; the whole verdict vector
    is checked through its colliding fraction against the sample's and through the host-buffer entry point."""
    import torch
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    from tests.common import cfg3_map
    side, npose = 8192, 10_000_000
    occ, _ = cfg3_map(side, seed=2)
    m = hpp.Map(ctx, occ, 0.2)
    om = O.Map(occ, 0.2)
    _, dist = m.layers()
    assert bits_equal(dist, om.distance)
    poses = random_poses(npose, side, side, margin=2.0, seed=3)
    d_poses = torch.from_numpy(poses).cuda()
    verd = torch.empty(npose, dtype=torch.uint8, device="cuda")
    m.collide_circles_dev(d_poses, npose, verd)
    ctx.sync()
    got = verd.cpu().numpy()
    want = O.collide_circles(om, poses[::100])
    assert np.array_equal(got[::100], want), f"{(got[::100] != want).sum()} of {len(want)} sampled verdicts differ"
    assert abs(got.mean() - want.mean()) < 5e-3
    assert np.array_equal(m.collide_circles(poses[:1_000_000]), got[:1_000_000])   # host-buffer path, same verdicts
    m.close()


def test_collide_compact_layer_equals_float_layer(ctx):
    """K4 reads the compact squared-distance layer (one byte per cell, clamped to 255, 8x4 micro-tiles) by default and
    the float layer for circle radii >= 14 cells: the verdicts must be identical on dense and on sparse maps (clamped
    taps), at the map border (index aliasing) and for a fine resolution where the radius is close to the limit."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    sparse = np.zeros((300, 333), np.int8)
    sparse[150:153, 100:104] = 100
    sparse[10, 300] = 100
    for occ, res in ((block_map(301, 203, 0.1, 2, 3), 0.2), (sparse, 0.2), (block_map(256, 256, 0.02, 4, 2), 0.05),
                     (np.zeros((40, 50), np.int8), 0.2)):
        H, W = occ.shape
        m = hpp.Map(ctx, occ, res)
        om = O.Map(occ, res)
        poses = random_poses(30000, W, H, res=res, margin=-2.0 * res * 10, seed=7)      # also outside the map
        got = m.collide_circles(poses)
        want = O.collide_circles(om, poses)
        assert np.array_equal(got, want), f"{(got != want).sum()} verdicts differ ({W}x{H} at {res})"
        m.close()


def test_collide_improved_matches_oracle(ctx):
    """Grid_map::isSingleStateCollisionFreeImproved (RS/src/Grid_map.cpp:286-317): all three outcomes occur"""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = block_map(200, 160, 0.01, 3, 4)
    m = hpp.Map(ctx, occ, 0.2)
    om = O.Map(occ, 0.2)
    poses = random_poses(20000, 200, 160, margin=-3.0, seed=9)
    got = m.collide_improved(poses)
    want = O.collide_improved(om, poses)
    assert np.array_equal(got, want)
    plain = O.collide_circles(om, poses)
    assert (want != plain).any() and (want == 0).any() and (want == 1).any()
    m.close()


@pytest.mark.parametrize("W,H,res,case", [(301, 203, 0.2, "blocks"), (90, 80, 0.2, "border"), (300, 7, 0.2, "narrow"),
                                          (160, 120, 1.0, "coarse")])
def test_collide_filter_path_edge_cases(ctx, W, H, res, case):
    """Batches of >= 16 384 poses take the float32 filter + exact list kernels (collide.cuh, circles_filter_u8): the verdicts
    must be the oracle's on poses the filter may not decide itself -- circles straddling the border, poses outside the
    map, headings beyond its range-reduction limit (|h| > 16), non-finite poses, a map too narrow for its border rule --
    as on ordinary ones."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    rng = np.random.default_rng(W + H)
    if case == "border":
        occ = np.zeros((H, W), np.int8)
        occ[H // 2, W // 2] = 100
    else:
        occ = block_map(W, H, 0.06, 3, seed=W)
    m = hpp.Map(ctx, occ, res)
    om = O.Map(occ, res)
    n = 40000
    ex, ey = W * om.res / 2, H * om.res / 2
    poses = np.empty((n, 3))
    poses[:, 0] = rng.uniform(-ex - 3.0, ex + 3.0, n)          # a band of 3 m around the map is outside
    poses[:, 1] = rng.uniform(-ey - 3.0, ey + 3.0, n)
    poses[:, 2] = rng.uniform(-np.pi, np.pi, n)
    k = n // 4                                                  # a quarter hugs the four borders
    side = rng.integers(0, 4, k)
    d = rng.uniform(-1.5, 1.5, k)
    t = rng.uniform(-1, 1, k)
    poses[:k, 0] = np.where(side == 0, -ex + d, np.where(side == 1, ex + d, t * ex))
    poses[:k, 1] = np.where(side == 2, -ey + d, np.where(side == 3, ey + d, t * ey))
    poses[k:k + 2000, 2] = rng.uniform(-40.0, 40.0, 2000)       # headings the filter hands to the exact path
    poses[k + 2000:k + 2100, 2] = rng.choice([16.0, -16.0, 15.999999, 16.000001, np.pi, -np.pi], 100)
    finite = np.ones(n, bool)
    bad = rng.choice(np.arange(k + 2100, n), 60, replace=False)
    poses[bad[:20], 0] = np.nan
    poses[bad[20:30], 1] = np.inf
    poses[bad[30:40], 0] = -np.inf
    poses[bad[40:50], 2] = np.nan
    poses[bad[50:60], 2] = np.inf
    finite[bad] = False
    got = m.collide_circles(poses)
    want = O.collide_circles(om, poses[finite])
    assert np.array_equal(got[finite], want), f"{(got[finite] != want).sum()} verdicts differ ({case})"
    assert case == "narrow" or 0.05 < want.mean() < 0.999   # (a 1.4 m wide map holds no free pose of a 1.9 m wide car)
    # non-finite poses: the same verdict from the filter path as from the one-pass exact kernel (small batches)
    small = m.collide_circles(poses[bad])
    assert np.array_equal(got[bad], small)
    m.close()
