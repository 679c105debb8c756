"""GPU parity: K3a/K3b heuristic fields through the C ABI vs the CPU oracle (bit-exact fp64)."""
import numpy as np
import pytest

from tests.common import block_map, bits_equal, cfg3_map

pytestmark = pytest.mark.gpu


def _goal_clear(occ, gx, gy):
    occ = occ.copy()
    occ[max(0, gy - 1):gy + 2, max(0, gx - 1):gx + 2] = 0
    return occ


@pytest.mark.parametrize("W,H,p,blk,seed", [(50, 50, 0.1, 1, 1), (120, 90, 0.3, 4, 2), (200, 200, 0.2, 1, 3),
                                            (300, 300, 0.3, 8, 4), (64, 64, 0.0, 1, 5), (65, 127, 0.15, 2, 6),
                                            (701, 710, 0.1, 3, 7)])
def test_heuristic_rs_bit_exact(ctx, W, H, p, blk, seed):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    gx, gy = W // 2, H // 3
    occ = _goal_clear(block_map(W, H, p, blk, seed), gx, gy)
    occ[0, 0] = 100
    m = hpp.Map(ctx, occ, 0.2)
    got, rep = m.heuristic_rs((gx, gy), (0, 0))
    want, q = O.heuristic_rs(occ, (gx, gy), (0, 0), quirks=True)
    if not (q.h1_fired or q.h2_clamped):
        assert bits_equal(got, want), f"{(got != want).sum()} cells differ"
    exact, _ = O.heuristic_rs(occ, (gx, gy), (0, 0), quirks=False)
    assert bits_equal(got, exact)
    assert rep.reachable == int(np.isfinite(exact).sum())
    assert rep.max_cost == exact[np.isfinite(exact)].max()
    m.close()


@pytest.mark.parametrize("W,H,p,blk,seed", [(50, 50, 0.1, 1, 1), (120, 90, 0.3, 4, 2), (200, 200, 0.2, 1, 3),
                                            (300, 300, 0.3, 8, 4), (65, 127, 0.15, 2, 6)])
def test_heuristic_dubins_bit_exact(ctx, W, H, p, blk, seed):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    gx, gy = W // 2, H // 3
    occ = _goal_clear(block_map(W, H, p, blk, seed), gx, gy)
    m = hpp.Map(ctx, occ, 0.2)
    got, rep = m.heuristic_dubins((gx, gy))
    want = O.heuristic_dubins(occ, (gx, gy))
    assert bits_equal(got, want), f"{(got != want).sum()} cells differ"
    m.close()


def test_heuristic_goal_on_obstacle_and_corner(ctx):
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = block_map(100, 80, 0.2, 2, 3)
    occ[10, 10] = 100
    m = hpp.Map(ctx, occ, 0.2)
    for goal in ((10, 10), (0, 0), (99, 79), (63, 63), (64, 64)):
        got, _ = m.heuristic_rs(goal, (-5, -5))
        assert bits_equal(got, O.heuristic_rs(occ, goal, (0, 0), quirks=False)[0])
        got, _ = m.heuristic_dubins(goal)
        assert bits_equal(got, O.heuristic_dubins(occ, goal))
    m.close()


def test_heuristic_quirk_report(ctx):
    """On a map where the reference's progress check (H1) fires the device reports it instead of claiming parity."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = block_map(400, 400, 0.05, 2, 7)
    occ[200, 200] = 0
    occ[10, 10] = 0
    m = hpp.Map(ctx, occ, 0.2)
    got, rep = m.heuristic_rs((200, 200), (10, 10))
    ref, q = O.heuristic_rs(occ, (200, 200), (10, 10), quirks=True)
    exact, _ = O.heuristic_rs(occ, (200, 200), (10, 10), quirks=False)
    assert q.h1_fired and (ref != exact).any()
    assert rep.h1_progress_break in (1, 2)
    assert bits_equal(got, exact)
    # start out of bounds: fires for certain once 70000 cells are settled
    _, rep2 = m.heuristic_rs((200, 200), (-1, -1))
    assert rep2.h1_progress_break == 1
    m.close()


def test_heuristic_bad_goal_is_an_error(ctx):
    import hybrid_path_planning_b200 as hpp
    m = hpp.Map(ctx, np.zeros((20, 20), np.int8), 0.2)
    with pytest.raises(hpp.HppError):
        m.heuristic_rs((20, 3), (0, 0))
    m.close()


def test_heuristic_cfg3_1024(ctx):
    """cfg3 recipe at 1024^2 (the oracle takes ~0.2 s): bit-exact, and the reference's quirks provably idle."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ, goal = cfg3_map(1024, seed=1)
    m = hpp.Map(ctx, occ, 0.2)
    got, rep = m.heuristic_rs(goal, (0, 0))
    want, q = O.heuristic_rs(occ, goal, (0, 0), quirks=True)
    assert not q.h1_fired and not q.h2_clamped
    assert rep.h1_progress_break == 0 and rep.h2_bucket_clamp == 0
    assert bits_equal(got, want)
    m.close()


def test_heuristic_cfg3_full_size(ctx):
    """BASELINE cfg3: 4096^2.  The oracle (reference-equivalent, ~3 s) is run in full: bit-exact field."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ, goal = cfg3_map(4096, seed=1)
    m = hpp.Map(ctx, occ, 0.2)
    got, rep = m.heuristic_rs(goal, (0, 0))
    want, q = O.heuristic_rs(occ, goal, (0, 0), quirks=True)
    assert not q.h1_fired and not q.h2_clamped
    assert rep.h1_progress_break == 0 and rep.h2_bucket_clamp == 0
    assert bits_equal(got, want), f"{(got != want).sum()} cells differ"
    m.close()


def test_heuristic_cfg3_bench_launch_148_fields(ctx):
    """The launch bench.py times: hpp_heuristic_batch_dev, 148 goal fields on the 4096^2 cfg3 map, default
    environment (one CTA of 20 warps per field, list schedule, longest-first ordering).  Fields of the patch goal, of
    two drawn goals and of the last goal are bit-compared with the oracle (reference-equivalent: quirks on, and the
    test asserts they do not fire); every other field is checked through size-independent properties: zero at its
    goal and nowhere else, +inf exactly on the occupied cells and the cells the oracle finds unreachable, and the
    fixed-point equation on a strided sample of rows."""
    import torch
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    from bench import patch_goals
    side, F = 4096, 148
    occ, _ = cfg3_map(side, seed=1)
    goals = np.array(patch_goals(side, F), np.int32)
    d_occ = torch.from_numpy(occ).cuda()
    m = hpp.Map(ctx, (side, side), 0.2, device_occ=d_occ)
    cost = torch.empty(F * side * side, dtype=torch.float64, device="cuda")
    reps = m.heuristic_batch_dev(hpp.VARIANT_RS, goals, cost, want_reports=True)
    ctx.sync()
    cost = cost.view(F, side, side)
    want0 = None
    for f in (0, 25, 77, F - 1):
        gx, gy = (int(v) for v in goals[f])
        want, q = O.heuristic_rs(occ, (gx, gy), (0, 0), quirks=True)
        assert not q.h1_fired and not q.h2_clamped
        got = cost[f].cpu().numpy()
        assert bits_equal(got, want), f"field {f}: {(got != want).sum()} cells differ"
        assert reps[f].reachable == int(np.isfinite(want).sum())
        if want0 is None:
            want0 = want
    # all fields: the reachable set is one connected component of the free cells, shared by every goal inside it
    reach = torch.from_numpy(np.isfinite(want0)).cuda()
    occ_t = d_occ.view(side, side) > 0
    for f in range(F):
        gx, gy = (int(v) for v in goals[f])
        c = cost[f]
        assert float(c[gy, gx]) == 0.0
        if bool(reach[gy, gx]):
            assert bool((torch.isfinite(c) == reach).all()), f
            assert int((c == 0).sum()) == 1
        assert bool(torch.isinf(c[occ_t]).all())   # every goal of the list is a free cell of the seed-1 map
    # fixed-point equation on rows 1, 65, 129, ... of every 16th field (fp64, the reference's two additions)
    near = torch.zeros_like(occ_t)
    near[1:, :] |= occ_t[:-1, :]
    near[:-1, :] |= occ_t[1:, :]
    near[:, 1:] |= occ_t[:, :-1]
    near[:, :-1] |= occ_t[:, 1:]
    rows = torch.arange(1, side - 1, 64, device="cuda")
    for f in range(0, F, 16):
        c = cost[f]
        best = torch.full((len(rows), side - 2), float("inf"), dtype=torch.float64, device="cuda")
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == 0 and dy == 0:
                    continue
                w = 1.0 if (dx == 0 or dy == 0) else 1.414
                wv = torch.where(near[rows, 1:-1], torch.tensor(w + 5.0, dtype=torch.float64, device="cuda"),
                                 torch.tensor(w, dtype=torch.float64, device="cuda"))
                best = torch.minimum(best, c[rows + dy][:, 1 + dx:side - 1 + dx] + wv)
        cur = c[rows, 1:-1]
        free = ~occ_t[rows, 1:-1]
        gx, gy = (int(v) for v in goals[f])
        isgoal = (rows[:, None] == gy) & (torch.arange(1, side - 1, device="cuda")[None, :] == gx)
        sel = free & ~isgoal
        assert bool((cur[sel] == best[sel]).all()), f"field {f}: fixed-point equation violated"
    m.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_heuristic_batch_host_chunks(ctx, variant):
    """hpp_heuristic_batch: 70 goal cells in one call (>= 64 fields are solved in two overlapped chunks), every field
    bit-identical to the oracle's fixed point; the reports carry the per-field reachable counts."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    occ = block_map(200, 150, 0.12, 3, seed=21)
    rng = np.random.default_rng(5)
    free = np.argwhere(occ == 0)
    goals = np.array([free[i][::-1] for i in rng.choice(len(free), 70, replace=False)], np.int32)
    goals[0] = (0, 0)
    goals[1] = (199, 149)
    m = hpp.Map(ctx, occ, 0.2)
    got, reps = m.heuristic_batch(variant, goals, want_reports=True)
    assert got.shape == (70, 150, 200)
    for f in (0, 1, 2, 17, 34, 35, 36, 69):
        g = tuple(int(v) for v in goals[f])
        want = O.heuristic_rs(occ, g, (0, 0), quirks=False)[0] if variant == 0 else O.heuristic_dubins(occ, g)
        assert bits_equal(got[f], want), f"field {f}: {(got[f] != want).sum()} cells differ"
        fin = want < (np.inf if variant == 0 else 1000.0)
        assert reps[f].reachable == int(fin.sum())
    m.close()


def test_heuristic_window_widths_and_staging_paths_agree(ctx, monkeypatch):
    """narrow and unbounded cost windows give the same bits on even- and odd-width maps (odd widths take the plain-load
    staging path, even ones the TMA tensor box copy), RS and Dubins, and on a 1024^2 cfg3-style map"""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    for W, H in ((330, 257), (331, 120)):
        occ = block_map(W, H, 0.25, 5, seed=W)
        occ[H // 2, W // 2] = 0
        want = O.heuristic_rs(occ, (W // 2, H // 2), (0, 0), quirks=False)[0]
        wantd = O.heuristic_dubins(occ, (W // 2, H // 2))
        m = hpp.Map(ctx, occ, 0.2)
        for delta in ("24", "1000000"):
            monkeypatch.setenv("HPP_HEUR_DELTA", delta)
            got, _ = m.heuristic_rs((W // 2, H // 2), (0, 0))
            assert bits_equal(got, want), (W, H, delta)
            gotd, _ = m.heuristic_dubins((W // 2, H // 2))
            assert bits_equal(gotd, wantd), (W, H, delta)
        m.close()
    monkeypatch.delenv("HPP_HEUR_DELTA")
    occ, goal = cfg3_map(1024, seed=3)
    want = O.heuristic_rs(occ, goal, (0, 0), quirks=False)[0]
    m = hpp.Map(ctx, occ, 0.2)
    got, rep = m.heuristic_rs(goal, (0, 0))
    assert bits_equal(got, want) and rep.reachable == int(np.isfinite(want).sum())
    m.close()


def test_heuristic_one_cta_per_field_list_schedule(ctx, monkeypatch):
    """HPP_HEUR_GROUP=1 forces the one-CTA-per-field mode (plain launch, list-based schedule of the rounds) that large
    batches use; checked here on single fields of several shapes and on a 150-field batch (RS and Dubins)."""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    monkeypatch.setenv("HPP_HEUR_GROUP", "1")
    for W, H, p, blk, seed in ((64, 64, 0.1, 2, 1), (301, 203, 0.25, 4, 2), (512, 384, 0.3, 8, 3)):
        occ = block_map(W, H, p, blk, seed)
        occ[H // 2, W // 2] = 0
        m = hpp.Map(ctx, occ, 0.2)
        got, _ = m.heuristic_rs((W // 2, H // 2), (0, 0))
        assert bits_equal(got, O.heuristic_rs(occ, (W // 2, H // 2), (0, 0), quirks=False)[0]), (W, H)
        gotd, _ = m.heuristic_dubins((W // 2, H // 2))
        assert bits_equal(gotd, O.heuristic_dubins(occ, (W // 2, H // 2))), (W, H)
        m.close()
    monkeypatch.delenv("HPP_HEUR_GROUP")
    occ = block_map(160, 128, 0.15, 3, seed=9)
    free = np.argwhere(occ == 0)
    rng = np.random.default_rng(2)
    goals = np.array([free[i][::-1] for i in rng.choice(len(free), 150, replace=False)], np.int32)
    m = hpp.Map(ctx, occ, 0.2)
    for variant in (0, 1):
        got = m.heuristic_batch(variant, goals)          # 150 fields: chunks of 38 -> groups; force the solo path below
        for f in (0, 37, 38, 149):
            g = tuple(int(v) for v in goals[f])
            want = O.heuristic_rs(occ, g, (0, 0), quirks=False)[0] if variant == 0 else O.heuristic_dubins(occ, g)
            assert bits_equal(got[f], want), (variant, f)
    import torch
    cost = torch.empty((150, 128, 160), dtype=torch.float64, device="cuda:0")
    m.heuristic_batch_dev(hpp.VARIANT_RS, goals, cost)    # 150 fields in one launch: one CTA per field
    ctx.sync()
    for f in (0, 75, 149):
        g = tuple(int(v) for v in goals[f])
        assert bits_equal(cost[f].cpu().numpy(), O.heuristic_rs(occ, g, (0, 0), quirks=False)[0]), f
    m.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_heuristic_subwindows_staging_and_cta_shapes_agree(ctx, monkeypatch, variant):
    """every sub-window width, the row-wise bulk-copy staging and the 4-warp CTAs that big batches on small maps use
    (HPP_HEUR_WARPS=4) give the oracle's bits -- on a map with long straight fronts (empty), thin diagonal ones (blocks)
    and an odd width (no TMA staging)"""
    import hybrid_path_planning_b200 as hpp
    from oracle import pyorc as O
    cases = [(np.zeros((96, 130), np.int8), (5, 90)), (block_map(161, 131, 0.2, 4, seed=7), (80, 65)),
             (block_map(224, 192, 0.3, 8, seed=8), (3, 3))]
    for occ, goal in cases:
        occ = occ.copy()
        occ[goal[1], goal[0]] = 0
        want = O.heuristic_rs(occ, goal, (0, 0), quirks=False)[0] if variant == 0 else O.heuristic_dubins(occ, goal)
        m = hpp.Map(ctx, occ, 0.2)
        for din in ("4", "32", "1000"):
            monkeypatch.setenv("HPP_HEUR_DELTA_IN", din)
            got = m.heuristic_rs(goal, (0, 0))[0] if variant == 0 else m.heuristic_dubins(goal)[0]
            assert bits_equal(got, want), (occ.shape, din)
        monkeypatch.delenv("HPP_HEUR_DELTA_IN")
        monkeypatch.setenv("HPP_HEUR_TMA2D", "0")       # row-wise bulk copies instead of one tensor box copy per tile
        got = m.heuristic_rs(goal, (0, 0))[0] if variant == 0 else m.heuristic_dubins(goal)[0]
        assert bits_equal(got, want), (occ.shape, "row-wise TMA staging")
        monkeypatch.delenv("HPP_HEUR_TMA2D")
        monkeypatch.setenv("HPP_HEUR_WARPS", "4")
        monkeypatch.setenv("HPP_HEUR_GROUP", "1")
        got = m.heuristic_rs(goal, (0, 0))[0] if variant == 0 else m.heuristic_dubins(goal)[0]
        assert bits_equal(got, want), (occ.shape, "4 warps")
        monkeypatch.delenv("HPP_HEUR_WARPS")
        monkeypatch.delenv("HPP_HEUR_GROUP")
        m.close()
