#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: "heuristic Mcells/s ... vs host CPU").

Workload at N=1 (config.workload = "cfg3"): BASELINE.json configs[2], the configuration the north_star's target is
quoted on -- "holonomic heuristic plus distance-layer EDT on a synthetic 4096x4096 random-obstacle grid".
One step = one pass of the hot path over one map message and a batch of F queries on it (--fields, default 148 = one per SM; the
reference arm runs the same kind of batch, one field per host core):  K1 (per-cell tables + obstacle/distance layers of the
occupancy, what the reference's Grid_map constructor does per map message) + K3a (the RS holonomic-with-obstacles
heuristic field of each query's goal, what HybridAstar::run does first).  value = F x 4096^2 cells x steps / device
time, inputs resident in HBM.  e2e = the same step through the host-buffer C ABI (hpp_map_update +
hpp_heuristic_batch): occupancy H2D from pinned memory, F cost fields D2H, inside the timed region.  The latency of a
single field (F = 1, hpp_heuristic_rs) is reported under extra.single_field.

Extras on the same JSON line (rank 0, N=1): collision checks/s on cfg4 (8192^2 distance layer, 10^7 poses) and Hybrid
A* queries/s on a cfg5 sample, each with a bounded CPU sample beside it.

--impl reference times the reference's own CPU implementation of the same step (oracle/_ref/libref_rs.so = the
reference's HybridAstar.cpp compiled -O2, when present; else the C oracle port) on all host cores, one field per core.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from tests.common import cfg3_map, parking_map, random_poses  # noqa: E402

METRIC = "heuristic Mcells/s (cfg3: K1 distance layers + K3a RS holonomic heuristic field, 4096x4096 map)"
UNIT = "Mcells/s"
RES = 0.2
HEUR_BYTES_PER_CELL = 9      # 1 B occupancy read + 8 B fp64 cost write
EDT_BYTES_PER_CELL = 9       # 1 B read + 4 B obstacle + 4 B distance written
COLLIDE_BYTES_PER_POSE = 89  # 24 B pose + 1 B verdict + 4 circles x 4 taps x 4 B


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def patch_goals(side, k):
    """k goal cells: the 25 cells of the cleared 5x5 patch at the map centre (cfg3 recipe) first, then free cells
    drawn with a fixed seed from the central quarter of the map (seed-1 map; every rank uses the same cell list and
    clears nothing else, so a drawn cell may be occupied on another rank's map -- the goal expands even then,
    RS/src/HybridAstar.cpp:378)"""
    c = side // 2
    g = [(c - 2 + (i % 5), c - 2 + ((i // 5) % 5)) for i in range(min(k, 25))]
    if k > 25:
        occ, _ = cfg3_map(side, seed=1)
        rng = np.random.default_rng(12345)
        while len(g) < k:
            x, y = (int(v) for v in rng.integers(side // 4, 3 * side // 4, 2))
            if occ[y, x] == 0 and (x, y) not in g:
                g.append((x, y))
    return g


def bench_config(side, F):
    """config of the headline line -- shared by both arms (the reference arm runs a bounded sample of this workload)"""
    return {"workload": "cfg3", "map": f"{side}x{side} 8x8-block Bernoulli(0.30), seed 1+rank", "resolution": RES,
            "fields_per_step": F,
            "goals": "the 25 cells of the cleared 5x5 goal patch, then seeded free cells of the central quarter",
            "step": f"one map message (layers: K1 / Grid_map ctor) + {F} goal fields on it "
                    f"(K3a / holonomicCostsWithObstacles_planning)",
            "per_rank": "own map",
            "l2": "flushed between steps by a 256 MiB write outside the per-step event pair; one cost field "
                  "(134 MB) alone exceeds the 126 MB L2",
            "timing": "CUDA events on the library's stream, per step, summed; max over ranks"}


# --------------------------------------------------------------------------------------------- reference arm
# The reference's own CPU implementation of the cfg3 step on all host cores.  Like the device arm, ONE layer build per
# map message and F goal fields on it:
#   * layers: cv2.distanceTransform (IPP off, all cores) -- the library call RS/src/Grid_map.cpp:64 makes -- built once
#     per map message in the parent and handed to the workers' planner instances through shared memory;
#   * fields: one worker process per core, each holding ONE planner instance per map message (reference's Grid_map
#     stand-in constructed from the supplied layers + HybridAstar, oracle/_ref) and calling the reference's
#     holonomicCostsWithObstacles_planning on it once per goal -- no second Grid_map, no layer rebuild per field.
# A timed "step" is a bounded sample of the workload: one wave of `cores` fields, plus cores/F of a layer build
# (the build is timed for real on every map message of the run and charged pro rata, exactly the device arm's ratio of
# one build per F fields).
_REF = {}  # parent -> forked workers: occupancy, shared layer buffers, side


def _ref_init():
    _REF["planner"] = None
    _REF["version"] = -1


def _ref_worker(task):
    """One goal field on one host core."""
    goal, version, want_sum = task
    from oracle import pyorc as O
    occ = _REF["occ"]
    if O.ref_available():
        if _REF["version"] != version:          # new map message: this worker's planner instance is rebuilt
            _REF["planner"] = None
            _REF["planner"] = O.Ref(0, occ, RES, layers=(_REF["obstacle"], _REF["distance"]))
            _REF["version"] = version
        h = _REF["planner"].heuristic(goal, want=want_sum)
    else:
        h, _ = O.heuristic_rs(occ, goal, (0, 0), True)
    if want_sum:
        return float(np.nan_to_num(h[goal[1], :], posinf=0.0).sum())
    return 0.0


def reference_arm(args):
    import multiprocessing as mp
    from oracle import pyorc as O
    O.build()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, args.ref_cores or 64)
    side, F = args.side, args.fields
    n = side * side
    goals = patch_goals(side, F)                # same goal list as the device arm
    kind = "reference" if O.ref_available() else "port"
    occ, _ = cfg3_map(side, seed=1)
    ctx = mp.get_context("fork")
    ob_raw, di_raw = ctx.RawArray("f", n), ctx.RawArray("f", n)
    _REF.update(occ=occ, obstacle=np.frombuffer(ob_raw, np.float32), distance=np.frombuffer(di_raw, np.float32))
    layer_ms = {}

    def map_message(threads):
        t0 = time.perf_counter()
        ob, di = O.cv2_layers(occ, RES, threads)
        _REF["obstacle"][:] = ob
        _REF["distance"][:] = di
        return (time.perf_counter() - t0) * 1e3

    layer_ms["1_thread"] = map_message(1)
    builds = [map_message(cores)]
    version = 0
    fields_done = 0
    with ctx.Pool(cores, initializer=_ref_init) as pool:
        def wave(k, want_sum=False):
            return pool.map(_ref_worker, [(goals[(k * cores + i) % F], version, want_sum) for i in range(cores)],
                            chunksize=1)
        for k in range(args.warmup):
            wave(k)
        t_waves = []
        sums = None
        for k in range(args.steps):
            if fields_done // F != (fields_done + cores) // F and k > 0:   # the next map message of the stream is due
                builds.append(map_message(cores))
                version += 1
            t0 = time.perf_counter()
            r = wave(k, want_sum=(k == 0))
            t_waves.append(time.perf_counter() - t0)
            fields_done += cores
            if k == 0:
                sums = r
    layer_ms["n_threads"] = min(builds)
    t_layer_share = layer_ms["n_threads"] / 1e3 * cores / F
    dt = sum(t_waves) + t_layer_share * args.steps
    value = n * cores * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(side, F),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"per step one wave of {cores} goal fields (one per core, first {cores}-field waves "
                                   f"of the device arm's {F}-goal list) by the reference's own "
                                   f"holonomicCostsWithObstacles_planning on one planner instance per worker and map "
                                   f"message, plus {cores}/{F} of one layer build (cv2.distanceTransform, IPP off, "
                                   f"{cores} threads: {layer_ms['n_threads']:.0f} ms; 1 thread: "
                                   f"{layer_ms['1_thread']:.0f} ms) -- the device arm's ratio of one layer build per "
                                   f"{F} fields; wall clock of the waves over the process pool, map synthesis excluded",
                         "layers_ms": layer_ms, "layer_builds_timed": len(builds),
                         "wave_ms": [t * 1e3 for t in t_waves[:8]],
                         "field_row_sums_first_wave": sums,
                         "flags": "-O2 -ffp-contract=off (upstream default is -O0)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- CPU legs of the extras
# (run in a child process started before/outside CUDA: fork after CUDA initialisation is unsafe)
def _cfg4_worker(task):
    lo, hi = task
    from oracle import pyorc as O
    return O.collide_circles(_REF["om"], _REF["poses"][lo:hi])


def _cfg5_worker(task):
    k, lo, hi = task
    from oracle import pyorc as O
    j = _REF["jobs"][k]
    use_ref = O.ref_available()
    pl = O.Ref(0, j["pm"], RES) if use_ref else O.Map(j["pm"], RES)
    solved = 0
    for i in range(lo, hi):
        if use_ref:
            solved += len(pl.run(tuple(j["starts"][i]), tuple(j["goals"][i]), max_states=20000)) > 1
        else:
            solved += len(O.run(pl, 0, tuple(j["starts"][i]), tuple(j["goals"][i]), quirks=True)["traj"]) > 1
    return solved


def cfg5_jobs(nmaps, npairs, collide):
    """cfg5 workload: nmaps synthetic parking maps x npairs collision-free start/goal pairs (collide(map k, poses) ->
    verdicts; device and oracle verdicts are bit-identical, so both arms draw the same pairs)"""
    jobs = []
    for k in range(nmaps):
        pm = parking_map(256, seed=k)
        cand = random_poses(6 * npairs + 2000, 256, 256, margin=3.0, seed=100 + k)
        free = cand[collide(k, pm, cand) == 0][:2 * npairs]
        jobs.append(dict(pm=pm, starts=free[:npairs], goals=free[npairs:2 * npairs]))
    return jobs


def cpu_extras(args):
    """cfg4 / cfg5 on all host cores: JSON on stdout"""
    import multiprocessing as mp
    from oracle import pyorc as O
    O.build()
    cores = min(os.cpu_count() or 1, args.ref_cores or 64)
    ctx = mp.get_context("fork")
    out = {}
    # cfg4: the full 8192^2 distance layer (cv2 EDT = the oracle's, bit-identical), a 2*10^6-pose sample of the 10^7
    side, ns = 8192, 2_000_000
    occ, _ = cfg3_map(side, seed=2)
    om = O.Map(occ, RES, with_layers=False)
    om.obstacle, om.distance = O.cv2_layers(occ, RES, cores)
    om.c.distance = om.distance.ctypes.data
    poses = random_poses(10_000_000, side, side, margin=2.0, seed=3)[:ns]
    _REF.update(om=om, poses=poses)
    with ctx.Pool(cores) as pool:
        chunks = [(i * ns // (4 * cores), (i + 1) * ns // (4 * cores)) for i in range(4 * cores)]
        pool.map(_cfg4_worker, chunks[:cores])
        t0 = time.perf_counter()
        v = np.concatenate(pool.map(_cfg4_worker, chunks, chunksize=1))
        t = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.collide_circles(om, poses[:200_000])
    t1 = time.perf_counter() - t0
    out["cfg4"] = {"checks_per_s": ns / t, "cores": cores, "kind": "port",
                   "sample": f"the first {ns} of the 10^7 poses on the full 8192^2 layer, C loop over the poses "
                             f"(orc_collide_circles_batch), {cores} worker processes; RS/src/Grid_map.cpp itself needs "
                             f"grid_map_core and cannot be compiled here, so this is the oracle port of "
                             f"isSingleStateCollisionFree",
                   "checks_per_s_1core": 200_000 / t1, "colliding_fraction": float(v.mean()),
                   "verdict_sum": int(v.sum())}
    del _REF["om"], _REF["poses"]
    # cfg5: all 8192 queries (8 maps x 1024 pairs) through the compiled reference's HybridAstar::run, one planner instance
    # per (worker, map chunk)
    nmaps, npairs = 8, 1024
    jobs = cfg5_jobs(nmaps, npairs, lambda k, pm, cand: O.collide_circles(O.Map(pm, RES), cand))
    _REF["jobs"] = jobs
    per = 64
    tasks = [(k, lo, min(lo + per, npairs)) for lo in range(0, npairs, per) for k in range(nmaps)]
    with ctx.Pool(cores) as pool:
        pool.map(_cfg5_worker, [(k, 0, 2) for k in range(min(nmaps, cores))])
        t0 = time.perf_counter()
        solved = sum(pool.map(_cfg5_worker, tasks, chunksize=1))
        t = time.perf_counter() - t0
    t0 = time.perf_counter()
    _cfg5_worker((0, 0, 64))
    t1 = time.perf_counter() - t0
    out["cfg5"] = {"queries_per_s": nmaps * npairs / t, "cores": cores,
                   "kind": "reference" if O.ref_available() else "port", "queries": nmaps * npairs, "solved": solved,
                   "sample": f"all {nmaps * npairs} queries, HybridAstar::run of the compiled reference, {cores} worker "
                             f"processes, {per} queries per task",
                   "queries_per_s_1core": 64 / t1}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples at or above the median (the sampler also sees the idle edges)
        load = [s for s in sm if s >= statistics.median(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- our arm
def ours_arm(args):
    import torch
    import hybrid_path_planning_b200 as hpp

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = hpp.Context(local_rank)          # raises when libhpp_b200.so or the GPU is missing: no fallback
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    side = args.side
    # weak scaling: every rank owns its own map (independent maps/goals shard with no data-path collective)
    occ, goal = cfg3_map(side, seed=1 + rank)
    n = side * side
    d_occ = torch.from_numpy(occ).cuda()
    m = hpp.Map(ctx, (side, side), RES, device_occ=d_occ)
    F = args.fields
    cost_dev = torch.empty(F * n, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    goals = np.array(patch_goals(side, F), np.int32)                   # distinct goal cells inside the cleared patch

    def step_dev():
        m.update_dev(d_occ)                                   # K1: cls + bit rows + obstacle/distance layers
        m.heuristic_batch_dev(hpp.VARIANT_RS, goals, cost_dev)  # K3a: F fields, one persistent launch

    for _ in range(args.warmup):
        step_dev()
    ctx.sync()
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    launches0 = ctx.launches
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            flush.fill_(k & 0xFF)                             # L2 flush, outside the per-step event pair
            ev[k][0].record(stream)
            m.update_dev(d_occ)
            ev[k][1].record(stream)
            m.heuristic_batch_dev(hpp.VARIANT_RS, goals, cost_dev)
            ev[k][2].record(stream)
    ctx.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    ms_steps = [e[0].elapsed_time(e[2]) for e in ev]
    ms_edt = [e[0].elapsed_time(e[1]) for e in ev]
    ms_heur = [e[1].elapsed_time(e[2]) for e in ev]
    t_dev = sum(ms_steps) / 1e3
    t_all = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    t_max = float(t_all.item())
    value = F * n * args.steps * world / t_max / 1e6

    # ---- e2e: the path a user of the planner runs, through the host-buffer C ABI.  Per step: one map message (occupancy
    # from pinned host memory: H2D + K1) and F start/goal queries on it -- hpp_plan_batch computes the F goal fields
    # (K3a, the same F cells as the device-timed step), keeps them ON THE DEVICE where the search consumes them, runs the
    # device-resident Hybrid A* and returns plan headers + trajectories to host memory.  Goal poses = centres of the F
    # goal cells, start poses = collision-free poses of this rank's map.
    res = float(np.float32(RES))
    org = -side * res / 2.0
    q_goals = np.stack([org + (goals[:, 0] + 0.5) * res, org + (goals[:, 1] + 0.5) * res, np.zeros(F)], 1)
    cand = random_poses(64 * F + 4000, side, side, margin=20.0, seed=5 + rank)
    q_starts = cand[m.collide_circles(cand) == 0][:F]
    assert len(q_starts) == F, "not enough collision-free start poses on the map"
    occ_pin = torch.from_numpy(occ).pin_memory()
    plans_pin = torch.zeros(F * hpp.capi.PLAN_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(hpp.capi.PLAN_DTYPE)
    traj_pin = torch.zeros((1 << 18, 5), dtype=torch.float64).pin_memory().numpy()
    occ_h = occ_pin.numpy()
    e2e_out = {}

    def e2e_step():
        m.update(occ_h)                                       # H2D occupancy + K1 + sync
        p, t = m.plan_batch(hpp.VARIANT_RS, q_starts, q_goals, plans_out=plans_pin, traj_out=traj_pin)  # fields + A* + D2H
        e2e_out["traj_rows"] = len(t)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    launches_e0 = ctx.launches
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        flush.fill_(1)
        e2e_step()
    torch.cuda.synchronize()
    barrier()
    t_e2e = time.perf_counter() - t0
    launches_e2e = (ctx.launches - launches_e0) // e2e_steps
    t_all = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    e2e_value = F * n * e2e_steps * world / float(t_all.item()) / 1e6
    e2e_h2d = int(n + 2 * F * 24)
    e2e_d2h = int(F * hpp.capi.PLAN_DTYPE.itemsize + e2e_out["traj_rows"] * 40 + 8)
    e2e_solved = int((plans_pin["n_states"] > 1).sum())
    e2e_iters = [float(plans_pin["iterations"].mean()), int(plans_pin["iterations"].max())]
    # sub-figure (rank 0 of a 1-GPU run): holonomicCostsWithObstacles_planning through host buffers -- the F fp64 fields
    # copied to pinned host memory (19.9 GB per step: PCIe-bound; the planner above never needs them on the host)
    fields_to_host = None
    cost_h = None
    if world == 1:
        import psutil
        whole = psutil.virtual_memory().available > 2.5 * F * n * 8
        Fe = F if whole else max(1, (F + 3) // 4)
        try:
            cost_pin = torch.empty(Fe * n, dtype=torch.float64, pin_memory=True)
        except RuntimeError:                                   # not enough lockable host memory: pageable buffer
            cost_pin = torch.empty(Fe * n, dtype=torch.float64)
        cost_h = cost_pin.numpy().reshape(Fe, side, side)

        def fields_step():
            m.update(occ_h)
            for f0 in range(0, F, Fe):
                m.heuristic_batch(hpp.VARIANT_RS, goals[f0:f0 + Fe], out=cost_h[:len(goals[f0:f0 + Fe])])
        fields_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            flush.fill_(1)
            fields_step()
        torch.cuda.synchronize()
        t_f = (time.perf_counter() - t0) / 2
        fields_to_host = {"value": F * n / t_f / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(n),
                          "d2h_bytes_per_step": int(8 * n * F), "fields_per_call": Fe,
                          "api": "hpp_map_update + hpp_heuristic_batch (host buffers, pinned)"}
    # latency of one field through the reference-shaped call (hpp_heuristic_rs = holonomicCostsWithObstacles_planning)
    one_h = cost_h[0] if cost_h is not None else torch.empty(n, dtype=torch.float64).pin_memory().numpy().reshape(side, side)
    m.heuristic_rs(goal, (0, 0), out=one_h)
    ts1 = []
    for _ in range(3):
        flush.fill_(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.heuristic_batch_dev(hpp.VARIANT_RS, goals[:1], cost_dev)
        ctx.sync()
        ts1.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    m.heuristic_rs(goal, (0, 0), out=one_h)
    t_single_e2e = time.perf_counter() - t0
    del cost_h, one_h

    # goal-row sums of the first fields of the device-resident batch: compared with the reference arm's (cpu_baseline)
    m.heuristic_batch_dev(hpp.VARIANT_RS, goals, cost_dev)
    ctx.sync()
    cview = cost_dev.view(F, side, side)
    row_sums = [float(np.nan_to_num(cview[f, int(goals[f][1]), :].cpu().numpy(), posinf=0.0).sum())
                for f in range(min(F, 64))]
    peak, peak_src = measured_peak()
    heur_ms = statistics.mean(ms_heur)
    achieved = HEUR_BYTES_PER_CELL * n * F / (heur_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("heuristic_dram_bytes_per_field")
            traffic = int(traffic * F) if traffic else None   # per launch, like `achieved` (ncu capture at 64 fields)
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(side, F),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                "steps": e2e_steps,
                "api": "hpp_map_update (occupancy from pinned host memory) + hpp_plan_batch (F start/goal queries: the F "
                       "goal fields stay on the device, device-resident Hybrid A*, plan headers + trajectories to host)",
                "queries_per_step": F, "solved": e2e_solved, "iterations_mean_max": e2e_iters,
                "gpu_launches_per_step": int(launches_e2e),
                "fields_to_host": fields_to_host},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": f"heur_warp_kernel (one persistent launch, {F} fields, one CTA of 20 warps per field, one warp per 32x32 tile, compacted sweeps)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes": HEUR_BYTES_PER_CELL * n * F,
                     "ms": heur_ms,
                     "k1_layers": {"ms": statistics.mean(ms_edt),
                                   "achieved": EDT_BYTES_PER_CELL * n / (statistics.mean(ms_edt) / 1e3) / 1e9}},
        "wall_s": t_wall,
    }

    line["extra"] = {"single_field": {"ms_device": min(ts1) * 1e3, "mcells_per_s_device": n / min(ts1) / 1e6,
                                      "ms_e2e": t_single_e2e * 1e3, "api": "hpp_heuristic_rs (host buffer, D2H inside)"}}
    if not args.no_sharded:
        line["sharded"] = sharded_section(ctx, torch, hpp, rank, world, dist if world > 1 else None, not args.quick)
    if rank == 0 and world == 1 and not args.no_extras:
        cpu_x = cpu_extras_child(args)
        line["extra"].update(extras(ctx, torch, hpp, stream, cpu_x, line.get("sharded", {}).get("cfg5")))
        cb = cpu_baseline(args)
        ref_sums = cb.pop("field_row_sums_first_wave", None) if isinstance(cb, dict) else None
        if ref_sums:
            cb["fields_checked_against_device"] = len(ref_sums)
            cb["device_fields_match_reference"] = bool(ref_sums == row_sums[:len(ref_sums)])
        line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.destroy_process_group()


def sharded_section(ctx, torch, hpp, rank, world, dist, full):
    """The partitioning the north_star names, through the C ABI's multi-GPU layer (hpp_comm: NCCL inside the library),
    at every N (N = 1 is the base of the scaling figures):
      cfg5      8192 RS queries = 8 parking maps x 1024 pairs, STRONG scaling: the query list is split contiguously
                over the ranks (maps replicated), every rank plans its slice with one hpp_plan_batch_multi call and the
                plan headers are gathered to every rank by one ncclAllGather (hpp_plan_batch_sharded);
                queries/s = 8192 / max-over-ranks wall time, warm-up pass first, best of 3.
      edt_band  distance layer of an 8192^2 and a 16384^2 map in row bands: pass 1 local, halo rows of the pass-1
                result exchanged with the neighbours (ncclSend/Recv), pass 2 (hpp_distance_layer_sharded); wall time
                with the exchange inside, max over ranks, best of 3."""
    dev = torch.device("cuda", torch.cuda.current_device())
    idt = torch.zeros(hpp.capi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(hpp.comm_unique_id()), dtype=torch.uint8))
    if world > 1:
        dist.broadcast(idt, 0)
    comm = hpp.Comm(ctx, idt.cpu().numpy().tobytes(), rank, world)

    def timed(fn, reps=3):
        best = float("inf")
        for _ in range(reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            fn()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best
    out = {}
    # ---- cfg5
    nmaps, npairs = 8, 1024
    maps = {}

    def dev_collide(k, pm, cand):                    # device verdicts (bit-exact with the oracle: tests)
        maps[k] = hpp.Map(ctx, pm, RES)
        return maps[k].collide_circles(cand)
    jobs = cfg5_jobs(nmaps, npairs, dev_collide)
    mlist = [maps[k] for k in range(nmaps)]
    starts = np.vstack([j["starts"] for j in jobs])
    goals = np.vstack([j["goals"] for j in jobs])
    moq = np.repeat(np.arange(nmaps, dtype=np.uint32), npairs)
    res = {}

    def run5():
        res["plans"], _, res["span"] = comm.plan_batch_sharded(mlist, moq, hpp.VARIANT_RS, starts, goals, traj_cap=1 << 21)
    run5()                                           # warm-up: device scratch grows once
    l0 = ctx.launches
    t5 = timed(run5)
    plans = res["plans"]
    out["cfg5"] = {"queries_per_s": nmaps * npairs / t5, "ms": t5 * 1e3, "queries": nmaps * npairs,
                   "solved": int((plans["n_states"] > 1).sum()), "scaling": "strong",
                   "queries_on_rank0": int(res["span"][1] - res["span"][0]),
                   "iterations_max": int(plans["iterations"].max()),
                   "gpu_launches_per_pass": int((ctx.launches - l0) // 3),
                   "api": "hpp_plan_batch_sharded (hpp_plan_batch_multi on the rank's slice of the query list: one call, "
                          "no host threads; ncclAllGather of the plan headers)"}
    for mm in mlist:
        mm.close()
    # ---- distance layer in row bands
    out["edt_band"] = {}
    for side in ((8192, 16384) if full else (8192,)):
        lo, hi = hpp.shard_range(side, world, rank)
        rng = np.random.default_rng(77)              # the same map on every rank, synthesised band by band
        b = rng.random((side // 8, side // 8)) < 0.30
        band_occ = torch.from_numpy(np.ascontiguousarray(
            (np.kron(b[lo // 8:(hi + 7) // 8], np.ones((8, 8), np.int8)) * 100).astype(np.int8)[lo - lo // 8 * 8:][:hi - lo])).to(dev)
        band = torch.empty((hi - lo) * side, dtype=torch.float32, device=dev)
        rep = {}

        def run_band():
            rep["r"] = comm.distance_layer_sharded(band_occ, side, side, RES, band, None, halo=64)
        run_band()
        tb = timed(run_band)
        r = rep["r"]
        out["edt_band"][f"{side}x{side}"] = {"ms": tb * 1e3, "rows_per_rank": int(hi - lo), "halo_rows": int(r.halo_rows),
                                             "rounds": int(r.rounds), "gathered_all": int(r.gathered_all),
                                             "halo_bytes_sent_rank0": int(r.halo_bytes_sent),
                                             "gcells_per_s": side * side / tb / 1e9,
                                             "band_checksum_rank0": float(band.double().sum().item())}
        del band_occ, band
    comm.close()
    return out


def k4_traffic():
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k4_collide_circles"]
        return int(d["dram_bytes_read"] + d["dram_bytes_write"])
    except Exception:
        return None


def extras(ctx, torch, hpp, stream, cpu_x, cfg5_dev):
    """cfg4 (collision checks/s) and cfg5 (queries/s): device numbers, with the all-core CPU figures of cpu_x beside."""
    from oracle import pyorc as O
    out = {}
    # ---- cfg4: 8192^2 distance layer, 10^7 poses
    side, npose = 8192, 10_000_000
    occ, _ = cfg3_map(side, seed=2)
    d_occ = torch.from_numpy(occ).cuda()
    m = hpp.Map(ctx, (side, side), RES, device_occ=d_occ)
    poses_h = random_poses(npose, side, side, margin=2.0, seed=3)
    poses = torch.from_numpy(poses_h).cuda()
    verd = torch.empty(npose, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        m.collide_circles_dev(poses, npose, verd)
    ctx.sync()
    ts = []
    with torch.cuda.stream(stream):
        for k in range(5):
            flush.fill_(k)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            m.collide_circles_dev(poses, npose, verd)
            b.record(stream)
            ctx.sync()
            ts.append(a.elapsed_time(b) / 1e3)
    t = statistics.mean(ts)
    peak, _ = measured_peak()
    pin = torch.from_numpy(poses_h).pin_memory().numpy()
    vout = torch.empty(npose, dtype=torch.uint8).pin_memory().numpy()
    m.collide_circles(pin, out=vout)
    t0 = time.perf_counter()
    m.collide_circles(pin, out=vout)
    t_e2e = time.perf_counter() - t0
    cpu4 = dict(cpu_x.get("cfg4", {"error": cpu_x.get("error", "no CPU leg")}))
    if "verdict_sum" in cpu4:   # the CPU leg checked the first 2*10^6 poses of the same list on the same map
        cpu4["device_verdicts_match"] = bool(int(verd[:2_000_000].sum().item()) == cpu4["verdict_sum"])
    out["cfg4_collision"] = {
        "checks_per_s": npose / t, "ms": t * 1e3, "colliding_fraction": float(verd.float().mean().item()),
        "roofline": {"bound": "hbm", "achieved": COLLIDE_BYTES_PER_POSE * npose / t / 1e9, "peak": peak,
                     "frac": COLLIDE_BYTES_PER_POSE * npose / t / 1e9 / peak, "unit": "GB/s",
                     "traffic": k4_traffic(),
                     "note": "round 2: (1) the taps read a compact squared-distance layer (one byte per cell, 8x4 micro-tiles = one "
                             "sector): measured DRAM traffic (ncu) 0.60 GB = 0.68x the algorithmic bytes (round 1: 3.22 GB = "
                             "3.6x at 87 % of the HBM peak); (2) a float32 filter with a proven certainty margin decides "
                             "99.97 % of the poses in 324 instead of 1 165 warp instructions per 32 poses, the exact fp64 "
                             "kernel works off the undecided rest (verdicts bit-identical): 0.50 -> 0.18 ms. 'achieved' "
                             "counts the algorithmic 89 B per pose (24 B pose + 1 B verdict + 16 float taps), which the "
                             "kernel no longer moves: the byte layer and the early exit make the physical traffic smaller"},
        "e2e_checks_per_s": npose / t_e2e, "h2d_bytes": 24 * npose, "d2h_bytes": npose,
        "cpu": cpu4}
    m.close()
    del poses, verd, d_occ
    try:
        out["cfg1_cfg2_latency"] = fixture_latency(ctx, hpp, O)
    except Exception as e:  # pragma: no cover  (never let an extra sink the headline line)
        out["cfg1_cfg2_latency"] = {"error": str(e)[-200:]}
    # ---- cfg5: the 1-GPU figure of the sharded section (one hpp_plan_batch_multi call for the 8 maps), CPU beside it
    out["cfg5_queries"] = dict(cfg5_dev or {"error": "sharded section skipped"})
    out["cfg5_queries"]["cpu"] = cpu_x.get("cfg5", {"error": cpu_x.get("error", "no CPU leg")})
    if "solved" in out["cfg5_queries"]["cpu"] and "solved" in out["cfg5_queries"]:
        out["cfg5_queries"]["cpu"]["device_solved_count_matches"] = bool(
            out["cfg5_queries"]["cpu"]["solved"] == out["cfg5_queries"]["solved"])
    return out


def fixture_latency(ctx, hpp, O):
    """BASELINE configs[0] / configs[1]: single start/goal queries on the reference's own fixture maps
    (tests/golden/maps.npz = the PNGs of grid_map_publisher in its OccupancyGrid convention), one hpp_plan_batch call
    per query (what the ROS node does), next to HybridAstar::run of the compiled reference (oracle/_ref; the oracle
    port when it is absent) on the same pairs, one host core, stopped after ~10 s per case.  Latency only."""
    maps = np.load(os.path.join(ROOT, "tests", "golden", "maps.npz"))
    res = float(np.float32(0.2))
    cases = [("cfg1 reeds_shepp_hybrid_astar, map_parking.png 0.2 m", maps["map_parking"], 0, 0.3, 12, None),
             ("cfg1b reeds_shepp_hybrid_astar, gridmap.png 0.2 m", maps["gridmap"], 0, 10.0, 6,
              ((66.8, -36.7, 0.785), (21.4, -45.7, 2.356))),
             ("cfg2 dubins_hybrid_star, map_maze.png 0.2 m inflated by 0.2 m", ctx.inflate(maps["map_maze"], res, 0.2), 1,
              0.5, 12, None)]
    out = {}
    for name, occ, variant, margin, n, extra_pair in cases:
        H, W = occ.shape
        t0 = time.perf_counter()
        m = hpp.Map(ctx, occ, res)
        ctx.sync()
        t_map = time.perf_counter() - t0
        cand = random_poses(4000, W, H, res=res, margin=margin, seed=11)
        v = m.collide_circles(cand) if variant == 0 else m.collide_sat(cand)
        free = cand[v == 0][:2 * n]
        k = len(free) // 2
        starts, goals = free[:k], free[k:2 * k]
        if extra_pair is not None:
            starts = np.vstack([starts, [extra_pair[0]]])
            goals = np.vstack([goals, [extra_pair[1]]])
        if variant == 0 and name.startswith("cfg1 "):
            starts = np.vstack([starts, [[-0.7, -4.9, np.pi]]])   # the solvable pair of SURVEY 8(d)
            goals = np.vstack([goals, [[-0.1, -4.7, np.pi]]])
        m.plan_batch(variant, starts[:1], goals[:1], traj_cap=1 << 16)
        ts, its, ok = [], [], 0
        for i in range(len(starts)):
            t0 = time.perf_counter()
            plans, _ = m.plan_batch(variant, starts[i:i + 1], goals[i:i + 1], traj_cap=1 << 16)
            ts.append(time.perf_counter() - t0)
            its.append(int(plans["iterations"][0]))
            ok += int(plans["n_states"][0] > 1)
        use_ref = O.ref_available()
        ref = O.Ref(variant, occ, res) if use_ref else None
        om = None if use_ref else O.Map(occ, res)
        tc, t_begin = [], time.perf_counter()
        for i in range(len(starts)):
            t0 = time.perf_counter()
            if use_ref:
                ref.run(tuple(starts[i]), tuple(goals[i]))
            else:
                O.run(om, variant, tuple(starts[i]), tuple(goals[i]), quirks=True)
            tc.append(time.perf_counter() - t0)
            if time.perf_counter() - t_begin > 10.0:
                break
        quirk = None
        if variant == 0 and W * H > 70000:
            # the reference's progress-check quirk H1 (RS/src/HybridAstar.cpp:452-472) fires on maps of more than 70 000
            # cells: count the heuristic cells on which the reference (oracle, quirks on) differs from the exact fixed
            # point the device returns, for the last query, next to what hpp_quirk_report says
            om_q = O.Map(occ, res, with_layers=False)
            st, gl = O._state(om_q, *starts[-1]), O._state(om_q, *goals[-1])
            gcell, scell = (int(gl.gridx), int(gl.gridy)), (int(st.gridx), int(st.gridy))
            hq, q = O.heuristic_rs(occ, gcell, scell, quirks=True)
            got, rep = m.heuristic_rs(gcell, scell)
            fin = np.isfinite(hq) & np.isfinite(got) & (got > 0)
            quirk = {"h1_fired_in_reference": bool(q.h1_fired), "cells_differing_from_reference": int((hq != got).sum()),
                     "max_rel_diff": float(np.max(np.abs(hq[fin] - got[fin]) / got[fin])) if fin.any() else 0.0,
                     "hpp_quirk_report_h1": int(rep.h1_progress_break),
                     "plans_equal_reference": "asserted in tests/test_gpu_baseline_configs.py::test_cfg1_rs_on_gridmap"}
        out[name] = {"cells": int(W * H), "queries": len(starts), "solved": ok, "iterations_mean": float(np.mean(its)),
                     "reference_quirk_h1": quirk,
                     "iterations_max": int(max(its)), "map_build_ms": t_map * 1e3,
                     "ms_per_query": 1e3 * float(np.mean(ts)), "ms_per_query_max": 1e3 * float(max(ts)),
                     "cpu": {"ms_per_query": 1e3 * float(np.mean(tc)), "ms_per_query_max": 1e3 * float(max(tc)),
                             "queries": len(tc), "kind": "reference" if use_ref else "port", "cores": 1,
                             "same_queries_device_ms": 1e3 * float(np.mean(ts[:len(tc)]))}}
        m.close()
    return out


def cpu_extras_child(args):
    """cfg4 / cfg5 on all host cores, in a child process (fork after CUDA init is unsafe)."""
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "cpu-extras"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env={**os.environ, "RANK": "0"})
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or "no output")[-300:]}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def cpu_baseline(args):
    """The reference arm on a bounded sample, in a child process (fork after CUDA init is unsafe)."""
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
               "--side", str(args.side), "--fields", str(args.fields)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0"})
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"value": None, "error": (r.stderr or "no output")[-300:]}
    except Exception as e:  # pragma: no cover
        return {"value": None, "error": str(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-extras"])
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--fields", type=int, default=148, help="goal fields (queries) per step; default = one per SM of a B200")
    ap.add_argument("--ref-cores", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the cfg5 / EDT-band section")
    ap.add_argument("--quick", action="store_true", help="skip the 16384^2 band case")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args)
    elif args.impl == "cpu-extras":
        cpu_extras(args)
    else:
        if args.warmup < 3:
            args.warmup = 3                 # timing rule: W >= 3
        ours_arm(args)


if __name__ == "__main__":
    main()
