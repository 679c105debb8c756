"""ctypes binding of include/hpp_b200.h (plumbing for tests and bench.py; not a compute path)."""
import ctypes as C
import weakref
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
VARIANT_RS, VARIANT_DUBINS = 0, 1
# GL/config/reeds_config.yaml:1-9 -- maxSteerAngle, wheelBase, axleToFront, axleToBack, width, pathLength, step_car
DEFAULT_PARAMS = (0.6, 2.75, 1.1, 1.2, 1.9, 4.0, 1.0)
MAX_COMMANDS, MAX_SIM = 26, 8


class HppError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("max_steer", "wheel_base", "axle_to_front", "axle_to_back", "width",
                                          "path_length", "step_car")]


class QuirkReport(C.Structure):
    _fields_ = [("reachable", C.c_uint64), ("max_cost", C.c_double), ("h1_progress_break", C.c_int32),
                ("h2_bucket_clamp", C.c_int32), ("rounds", C.c_uint32), ("tile_visits", C.c_uint32)]


class BandReport(C.Structure):
    _fields_ = [("rounds", C.c_uint32), ("halo_rows", C.c_uint32), ("gathered_all", C.c_uint32), ("pad_", C.c_uint32),
                ("halo_bytes_sent", C.c_uint64)]


class Node(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("heading", C.c_double), ("cost", C.c_double),
                ("steer", C.c_double), ("direction", C.c_int32), ("pad_", C.c_int32)]


class Child(C.Structure):
    _fields_ = [("traj", (C.c_double * 5) * MAX_SIM), ("cost", C.c_double), ("steer", C.c_double),
                ("direction", C.c_int32), ("valid", C.c_int32)]


class Shot(C.Structure):
    _fields_ = [("found", C.c_int32), ("n_paths", C.c_int32), ("chosen", C.c_int32), ("n_samples", C.c_int32),
                ("cost", C.c_double), ("sample_offset", C.c_uint64)]


class Plan(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("iterations", C.c_int32), ("via_shot", C.c_int32), ("status", C.c_int32),
                ("cost", C.c_double), ("traj_offset", C.c_uint64)]


NODE_DTYPE = np.dtype([("x", "f8"), ("y", "f8"), ("heading", "f8"), ("cost", "f8"), ("steer", "f8"),
                       ("direction", "i4"), ("pad_", "i4")])
CHILD_DTYPE = np.dtype([("traj", "f8", (MAX_SIM, 5)), ("cost", "f8"), ("steer", "f8"), ("direction", "i4"),
                        ("valid", "i4")])
SHOT_DTYPE = np.dtype([("found", "i4"), ("n_paths", "i4"), ("chosen", "i4"), ("n_samples", "i4"), ("cost", "f8"),
                       ("sample_offset", "u8")])
PLAN_DTYPE = np.dtype([("n_states", "i4"), ("iterations", "i4"), ("via_shot", "i4"), ("status", "i4"),
                       ("cost", "f8"), ("traj_offset", "u8")])

_lib = None


def lib_path():
    """HPP_B200_LIB: another build of the same library (A/B measurements of two builds inside one GPU session)"""
    return os.environ.get("HPP_B200_LIB") or os.path.join(HERE, "libhpp_b200.so")


def lib():
    """Load libhpp_b200.so.  Fails loudly when it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise HppError(f"{p} is missing: run hybrid_path_planning_b200/csrc/build.sh (or __graft_entry__.build())")
        L = C.CDLL(p)
        vp, i32, u32, u64, dbl, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_double, C.c_size_t
        sigs = {
            "hpp_ctx_create": [C.c_int, C.POINTER(vp)],
            "hpp_ctx_destroy": [vp],
            "hpp_ctx_sync": [vp],
            "hpp_map_create": [vp, vp, u32, u32, C.c_float, dbl, dbl, C.POINTER(Params), C.POINTER(vp)],
            "hpp_map_create_dev": [vp, vp, u32, u32, C.c_float, dbl, dbl, C.POINTER(Params), C.POINTER(vp)],
            "hpp_map_destroy": [vp],
            "hpp_map_set_params": [vp, C.POINTER(Params)],
            "hpp_map_update": [vp, vp],
            "hpp_map_update_dev": [vp, vp],
            "hpp_map_layers": [vp, vp, vp],
            "hpp_map_layers_dev": [vp, C.POINTER(vp), C.POINTER(vp)],
            "hpp_map_cells": [vp, vp, sz, vp, vp],
            "hpp_inflate": [vp, vp, u32, u32, dbl, dbl, vp],
            "hpp_inflate_dev": [vp, vp, u32, u32, dbl, dbl, vp],
            "hpp_cv_threshold_otsu": [vp, vp, u32, u32, vp, C.POINTER(dbl)],
            "hpp_cv_dilate_ellipse": [vp, vp, u32, u32, i32, i32, vp],
            "hpp_image_to_occupancy": [vp, vp, u32, u32, vp],
            "hpp_image_to_occupancy_dev": [vp, vp, u32, u32, vp],
            "hpp_collide_circles": [vp, vp, sz, vp],
            "hpp_collide_circles_dev": [vp, vp, sz, vp],
            "hpp_collide_sat": [vp, vp, sz, vp],
            "hpp_collide_sat_dev": [vp, vp, sz, vp],
            "hpp_collide_improved": [vp, vp, sz, vp],
            "hpp_collide_improved_dev": [vp, vp, sz, vp],
            "hpp_heuristic_rs": [vp, i32, i32, i32, i32, vp, C.POINTER(QuirkReport)],
            "hpp_heuristic_dubins": [vp, i32, i32, vp, C.POINTER(QuirkReport)],
            "hpp_heuristic_batch_dev": [vp, C.c_int, vp, sz, vp, vp],
            "hpp_heuristic_batch": [vp, C.c_int, vp, sz, vp, vp],
            "hpp_expand": [vp, C.c_int, vp, sz, vp],
            "hpp_expand_dev": [vp, C.c_int, vp, sz, vp],
            "hpp_rs_shot": [vp, vp, sz, vp, sz, vp, vp, sz, C.POINTER(sz)],
            "hpp_dubins_shot": [vp, vp, sz, vp, sz, vp, vp, sz, C.POINTER(sz)],
            "hpp_rs_all_paths": [vp, vp, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp],
            "hpp_test_heap_ops": [vp, vp, vp, sz, vp, vp],
            "hpp_dubins_path": [vp, vp, vp, dbl, dbl, vp, vp, vp, vp, sz],
            "hpp_plan_batch": [vp, C.c_int, vp, vp, sz, vp, vp, sz, C.POINTER(sz)],
            "hpp_plan_batch_multi": [vp, sz, vp, C.c_int, vp, vp, sz, vp, vp, sz, C.POINTER(sz)],
            "hpp_comm_unique_id": [vp],
            "hpp_comm_create": [vp, vp, C.c_int, C.c_int, C.POINTER(vp)],
            "hpp_plan_batch_sharded": [vp, vp, sz, vp, C.c_int, vp, vp, sz, vp, vp, sz, C.POINTER(sz), C.POINTER(sz),
                                       C.POINTER(sz)],
            "hpp_collide_circles_sharded": [vp, vp, vp, sz, vp],
            "hpp_distance_layer_sharded": [vp, vp, u32, u32, C.c_float, u32, vp, vp, C.POINTER(BandReport)],
        }
        for name, args in sigs.items():
            if hasattr(L, name):
                f = getattr(L, name)
                f.argtypes = args
                f.restype = C.c_int
        L.hpp_last_error.argtypes = [vp]
        L.hpp_last_error.restype = C.c_char_p
        L.hpp_ctx_stream.argtypes = [vp]
        L.hpp_ctx_stream.restype = vp
        L.hpp_ctx_launch_count.argtypes = [vp]
        L.hpp_ctx_launch_count.restype = u64
        L.hpp_version.restype = C.c_char_p
        L.hpp_comm_destroy.argtypes = [vp]
        L.hpp_comm_destroy.restype = None
        L.hpp_comm_rank.argtypes = L.hpp_comm_size.argtypes = [vp]
        L.hpp_comm_rank.restype = L.hpp_comm_size.restype = C.c_int
        L.hpp_shard_range.argtypes = [sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]
        L.hpp_shard_range.restype = None
        L.hpp_num_commands.argtypes = [C.POINTER(Params), C.c_int]
        L.hpp_num_commands.restype = i32
        _lib = L
    return _lib


def _ptr(a):
    """host numpy array or raw device address -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (device memory owned by torch: plumbing)
        return a.data_ptr()
    return int(a)


class Context:
    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        st = self.L.hpp_ctx_create(device, C.byref(h))
        if st != 0:
            raise HppError(f"hpp_ctx_create(device={device}) failed with status {st}: no usable sm_100 GPU "
                           "(there is no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, st):
        if st != 0:
            raise HppError(f"status {st}: {self.L.hpp_last_error(self.h).decode()}")

    def sync(self):
        self.check(self.L.hpp_ctx_sync(self.h))

    @property
    def stream(self):
        return self.L.hpp_ctx_stream(self.h)

    @property
    def launches(self):
        return int(self.L.hpp_ctx_launch_count(self.h))

    def inflate(self, occ, resolution, radius):
        occ = np.ascontiguousarray(occ, np.int8)
        H, W = occ.shape
        out = np.empty_like(occ)
        self.check(self.L.hpp_inflate(self.h, occ.ctypes.data, W, H, float(resolution), float(radius), out.ctypes.data))
        return out

    def cv_threshold_otsu(self, gray):
        """cv2.threshold(gray, 0, 255, THRESH_BINARY | THRESH_OTSU) -> (threshold, binary image)"""
        gray = np.ascontiguousarray(gray, np.uint8)
        H, W = gray.shape
        out = np.empty_like(gray)
        thr = C.c_double(0)
        self.check(self.L.hpp_cv_threshold_otsu(self.h, gray.ctypes.data, W, H, out.ctypes.data, C.byref(thr)))
        return thr.value, out

    def cv_dilate_ellipse(self, binary, r, complemented=False):
        """cv2.dilate(binary, getStructuringElement(MORPH_ELLIPSE, (2r+1, 2r+1), (r, r))) for a 0 / non-zero image;
        complemented=True: `binary` holds the complement of the image to dilate (a folded cv2.bitwise_not)"""
        binary = np.ascontiguousarray(binary, np.uint8)
        H, W = binary.shape
        out = np.empty_like(binary)
        self.check(self.L.hpp_cv_dilate_ellipse(self.h, binary.ctypes.data, W, H, int(r), int(bool(complemented)),
                                                out.ctypes.data))
        return out

    def image_to_occupancy(self, img):
        """8-bit gray image [rows, cols] -> occupancy [H = cols, W = rows] (grid_map_publisher's conversion)"""
        img = np.ascontiguousarray(img, np.uint8)
        rows, cols = img.shape
        occ = np.empty((cols, rows), np.int8)
        self.check(self.L.hpp_image_to_occupancy(self.h, img.ctypes.data, rows, cols, occ.ctypes.data))
        return occ

    def inflate_dev(self, occ_dev, W, H, resolution, radius, out_dev):
        self.check(self.L.hpp_inflate_dev(self.h, _ptr(occ_dev), W, H, float(resolution), float(radius), _ptr(out_dev)))

    def test_heap_ops(self, keys, ops):
        """unit hook: (payload of every pop, index of the first operation after which the warp-collective heap differs
        from the one-thread replay, -1 if never)"""
        keys = np.ascontiguousarray(keys, np.float64)
        ops = np.ascontiguousarray(ops, np.int32)
        pops = np.full(len(ops), -1, np.int32)
        bad = C.c_int64(0)
        self.check(self.L.hpp_test_heap_ops(self.h, keys.ctypes.data, ops.ctypes.data, len(ops), pops.ctypes.data,
                                            C.byref(bad)))
        return pops, bad.value

    def rs_all_paths(self, s, g, maxc, step=1.0, max_paths=48, samples_cap=400000):
        s3 = (C.c_double * 3)(*s)
        g3 = (C.c_double * 3)(*g)
        n_paths = C.c_int32(0)
        n_seg = np.zeros(max_paths, np.int32)
        lengths = np.zeros((max_paths, 5))
        types = np.zeros((max_paths, 5), np.uint8)
        tl = np.zeros(max_paths)
        ns = np.zeros(max_paths, np.int32)
        samples = np.zeros((samples_cap, 3))
        dirs = np.zeros(samples_cap, np.int32)
        self.check(self.L.hpp_rs_all_paths(self.h, s3, g3, maxc, step, max_paths, C.byref(n_paths), n_seg.ctypes.data,
                                           lengths.ctypes.data, types.ctypes.data, tl.ctypes.data, ns.ctypes.data,
                                           samples.ctypes.data, samples_cap, dirs.ctypes.data))
        out, off = [], 0
        for k in range(n_paths.value):
            m = int(ns[k])
            out.append(dict(lengths=lengths[k, :n_seg[k]].copy(), types=bytes(types[k, :n_seg[k]]).decode(), L=tl[k],
                            x=samples[off:off + m, 0].copy(), y=samples[off:off + m, 1].copy(),
                            yaw=samples[off:off + m, 2].copy(), dir=dirs[off:off + m].copy()))
            off += m
        return out

    def dubins_path(self, s, g, curv, step=0.1, samples_cap=100000):
        s3 = (C.c_double * 3)(*s)
        g3 = (C.c_double * 3)(*g)
        length = C.c_double(0)
        mode = C.create_string_buffer(4)
        ns = C.c_int32(0)
        samples = np.zeros((samples_cap, 3))
        self.check(self.L.hpp_dubins_path(self.h, s3, g3, curv, step, C.byref(length), mode, C.byref(ns),
                                          samples.ctypes.data, samples_cap))
        m = ns.value
        return dict(length=length.value, mode=mode.raw[:3].decode().rstrip("\0"), x=samples[:m, 0].copy(),
                    y=samples[:m, 1].copy(), yaw=samples[:m, 2].copy())

    def _adopt(self, child):
        """maps and communicators live inside the context: they are closed before it, whatever order the interpreter
        finalises the Python objects in (a map destroyed after its context would touch freed memory)"""
        if not hasattr(self, "_children"):
            self._children = weakref.WeakSet()
        self._children.add(child)

    def close(self):
        if getattr(self, "h", None):
            for child in list(getattr(self, "_children", ())):
                child.close()
            self.L.hpp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Map:
    """Device-resident Grid_map (occupancy + obstacle/distance layers + per-cell tables)."""

    def __init__(self, ctx, occ, resolution=0.2, origin=None, params=DEFAULT_PARAMS, device_occ=None):
        self.ctx, self.L = ctx, ctx.L
        self.params = Params(*params)
        res_f = np.float32(resolution)
        if device_occ is None:
            occ = np.ascontiguousarray(occ, np.int8)
            self.H, self.W = occ.shape
        else:
            self.W, self.H = occ  # (W, H) when the occupancy lives on the device
        if origin is None:
            origin = (-self.W * float(res_f) / 2.0, -self.H * float(res_f) / 2.0)
        self.res = float(res_f)
        self.ox, self.oy = float(origin[0]), float(origin[1])
        h = C.c_void_p()
        if device_occ is None:
            st = self.L.hpp_map_create(ctx.h, occ.ctypes.data, self.W, self.H, C.c_float(float(res_f)), self.ox, self.oy,
                                       C.byref(self.params), C.byref(h))
        else:
            st = self.L.hpp_map_create_dev(ctx.h, _ptr(device_occ), self.W, self.H, C.c_float(float(res_f)), self.ox,
                                           self.oy, C.byref(self.params), C.byref(h))
        ctx.check(st)
        self.h = h
        self.n = self.W * self.H
        ctx._adopt(self)

    def update(self, occ):
        occ = np.ascontiguousarray(occ, np.int8)
        self.ctx.check(self.L.hpp_map_update(self.h, occ.ctypes.data))

    def update_dev(self, occ_dev):
        self.ctx.check(self.L.hpp_map_update_dev(self.h, _ptr(occ_dev)))

    def layers(self):
        obstacle = np.empty(self.n, np.float32)
        distance = np.empty(self.n, np.float32)
        self.ctx.check(self.L.hpp_map_layers(self.h, obstacle.ctypes.data, distance.ctypes.data))
        return obstacle, distance

    def layers_dev(self):
        a, b = C.c_void_p(), C.c_void_p()
        self.ctx.check(self.L.hpp_map_layers_dev(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def cells(self, xy):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        cell = np.empty((len(xy), 2), np.int32)
        flags = np.empty(len(xy), np.uint8)
        self.ctx.check(self.L.hpp_map_cells(self.h, xy.ctypes.data, len(xy), cell.ctypes.data, flags.ctypes.data))
        return cell, flags

    def collide_circles(self, xyh, out=None):
        xyh = np.ascontiguousarray(xyh, np.float64).reshape(-1, 3)
        v = np.empty(len(xyh), np.uint8) if out is None else out
        self.ctx.check(self.L.hpp_collide_circles(self.h, xyh.ctypes.data, len(xyh), v.ctypes.data))
        return v

    def collide_sat(self, xyh):
        xyh = np.ascontiguousarray(xyh, np.float64).reshape(-1, 3)
        v = np.empty(len(xyh), np.uint8)
        self.ctx.check(self.L.hpp_collide_sat(self.h, xyh.ctypes.data, len(xyh), v.ctypes.data))
        return v

    def collide_improved(self, xyh):
        """Grid_map::isSingleStateCollisionFreeImproved, with the reference's own return convention (see the header)"""
        xyh = np.ascontiguousarray(xyh, np.float64).reshape(-1, 3)
        out = np.empty(len(xyh), np.uint8)
        self.ctx.check(self.L.hpp_collide_improved(self.h, xyh.ctypes.data, len(xyh), out.ctypes.data))
        return out

    def collide_circles_dev(self, xyh_dev, n, verdict_dev):
        self.ctx.check(self.L.hpp_collide_circles_dev(self.h, _ptr(xyh_dev), n, _ptr(verdict_dev)))

    def collide_sat_dev(self, xyh_dev, n, verdict_dev):
        self.ctx.check(self.L.hpp_collide_sat_dev(self.h, _ptr(xyh_dev), n, _ptr(verdict_dev)))

    def heuristic_rs(self, goal, start=(0, 0), out=None):
        cost = np.empty((self.H, self.W), np.float64) if out is None else out
        rep = QuirkReport()
        self.ctx.check(self.L.hpp_heuristic_rs(self.h, goal[0], goal[1], start[0], start[1], cost.ctypes.data,
                                               C.byref(rep)))
        return cost, rep

    def heuristic_dubins(self, goal):
        cost = np.empty((self.H, self.W), np.float64)
        rep = QuirkReport()
        self.ctx.check(self.L.hpp_heuristic_dubins(self.h, goal[0], goal[1], cost.ctypes.data, C.byref(rep)))
        return cost, rep

    def heuristic_batch_dev(self, variant, goals, cost_dev, want_reports=False):
        goals = np.ascontiguousarray(goals, np.int32).reshape(-1, 2)
        reps = (QuirkReport * len(goals))() if want_reports else None
        self.ctx.check(self.L.hpp_heuristic_batch_dev(self.h, variant, goals.ctypes.data, len(goals), _ptr(cost_dev),
                                                      C.cast(reps, C.c_void_p) if reps is not None else None))
        return reps

    def heuristic_batch(self, variant, goals, out=None, want_reports=False):
        """n goal cells -> n fields in a host buffer [n, H, W] (pass a pinned `out` to avoid the staging copy)"""
        goals = np.ascontiguousarray(goals, np.int32).reshape(-1, 2)
        cost = out if out is not None else np.empty((len(goals), self.H, self.W), np.float64)
        assert cost.dtype == np.float64 and cost.size == len(goals) * self.n and cost.flags.c_contiguous
        reps = (QuirkReport * len(goals))() if want_reports else None
        self.ctx.check(self.L.hpp_heuristic_batch(self.h, variant, goals.ctypes.data, len(goals), cost.ctypes.data,
                                                  C.cast(reps, C.c_void_p) if reps is not None else None))
        return (cost, reps) if want_reports else cost

    def num_commands(self, variant):
        return int(self.L.hpp_num_commands(C.byref(self.params), variant))

    def expand(self, variant, parents):
        """parents: structured array NODE_DTYPE -> children [n, n_cmd] CHILD_DTYPE"""
        parents = np.ascontiguousarray(parents, NODE_DTYPE)
        nc = self.num_commands(variant)
        kids = np.zeros((len(parents), nc), CHILD_DTYPE)
        self.ctx.check(self.L.hpp_expand(self.h, variant, parents.ctypes.data, len(parents), kids.ctypes.data))
        return kids

    def expand_dev(self, variant, parents_dev, n, children_dev):
        """device pointers: parents n x hpp_node, children n x n_cmd x hpp_child; enqueue only"""
        self.ctx.check(self.L.hpp_expand_dev(self.h, variant, _ptr(parents_dev), n, _ptr(children_dev)))

    def _shot(self, fn, starts, goals, samples_cap):
        starts = np.ascontiguousarray(starts, NODE_DTYPE)
        goals = np.ascontiguousarray(goals, np.float64).reshape(-1, 3)
        stride = 0 if len(goals) == 1 else 3
        shots = np.zeros(len(starts), SHOT_DTYPE)
        samples = np.zeros((samples_cap, 5))
        used = C.c_size_t(0)
        self.ctx.check(fn(self.h, starts.ctypes.data, len(starts), goals.ctypes.data, stride, shots.ctypes.data,
                          samples.ctypes.data, samples_cap, C.byref(used)))
        return shots, samples[:used.value]

    def rs_shot(self, starts, goals, samples_cap=1 << 20):
        return self._shot(self.L.hpp_rs_shot, starts, goals, samples_cap)

    def dubins_shot(self, starts, goals, samples_cap=1 << 20):
        return self._shot(self.L.hpp_dubins_shot, starts, goals, samples_cap)

    def plan_batch(self, variant, starts, goals, traj_cap=1 << 22, plans_out=None, traj_out=None):
        """plans_out / traj_out: caller-owned result buffers (e.g. pinned), otherwise fresh arrays"""
        starts = np.ascontiguousarray(starts, np.float64).reshape(-1, 3)
        goals = np.ascontiguousarray(goals, np.float64).reshape(-1, 3)
        plans = plans_out if plans_out is not None else np.zeros(len(starts), PLAN_DTYPE)
        traj = traj_out if traj_out is not None else np.empty((traj_cap, 5))
        assert len(plans) == len(starts) and plans.dtype == PLAN_DTYPE and traj.dtype == np.float64
        used = C.c_size_t(0)
        self.ctx.check(self.L.hpp_plan_batch(self.h, variant, starts.ctypes.data, goals.ctypes.data, len(starts),
                                             plans.ctypes.data, traj.ctypes.data, len(traj), C.byref(used)))
        return plans, traj[:used.value]

    def close(self):
        if getattr(self, "h", None):
            self.L.hpp_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


COMM_ID_BYTES = 128


def comm_unique_id():
    """128-byte NCCL id made by rank 0; ship it to the other ranks by any means (bench.py: torch.distributed)"""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    st = lib().hpp_comm_unique_id(C.cast(buf, C.c_void_p))
    if st != 0:
        raise HppError(f"hpp_comm_unique_id failed with status {st} (is libnccl.so.2 loadable?)")
    return bytes(buf)


def shard_range(n, world, rank):
    lo, hi = C.c_size_t(0), C.c_size_t(0)
    lib().hpp_shard_range(n, world, rank, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


class Comm:
    """hpp_comm: this rank's NCCL communicator inside the C library (one process per GPU)"""

    def __init__(self, ctx, unique_id, rank, world):
        self.L = lib()
        self.ctx = ctx
        self.rank, self.world = rank, world
        h = C.c_void_p()
        idb = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        ctx.check(self.L.hpp_comm_create(ctx.h, C.cast(idb, C.c_void_p), rank, world, C.byref(h)))
        self.h = h
        ctx._adopt(self)

    def plan_batch_sharded(self, maps, map_of_query, variant, starts, goals, traj_cap=1 << 22):
        """all ranks pass the whole query list; returns (plan headers of ALL queries, this rank's trajectories, (lo, hi))"""
        starts = np.ascontiguousarray(starts, np.float64).reshape(-1, 3)
        goals = np.ascontiguousarray(goals, np.float64).reshape(-1, 3)
        moq = None if map_of_query is None else np.ascontiguousarray(map_of_query, np.uint32).reshape(-1)
        handles = (C.c_void_p * len(maps))(*[m.h for m in maps])
        plans = np.zeros(len(starts), PLAN_DTYPE)
        traj = np.empty((traj_cap, 5))
        used, lo, hi = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        self.ctx.check(self.L.hpp_plan_batch_sharded(self.h, C.cast(handles, C.c_void_p), len(maps),
                                                     moq.ctypes.data if moq is not None else None, variant,
                                                     starts.ctypes.data, goals.ctypes.data, len(starts),
                                                     plans.ctypes.data, traj.ctypes.data, traj_cap, C.byref(used),
                                                     C.byref(lo), C.byref(hi)))
        return plans, traj[:used.value], (lo.value, hi.value)

    def collide_circles_sharded(self, m, xyh):
        xyh = np.ascontiguousarray(xyh, np.float64).reshape(-1, 3)
        out = np.empty(len(xyh), np.uint8)
        self.ctx.check(self.L.hpp_collide_circles_sharded(self.h, m.h, xyh.ctypes.data, len(xyh), out.ctypes.data))
        return out

    def distance_layer_sharded(self, occ_band_dev, W, total_rows, resolution, band_out_dev, full_out_dev=None, halo=0):
        """occ_band_dev: this rank's rows [lo, hi) of the occupancy (device); band_out_dev: (hi-lo) x W floats (device)"""
        rep = BandReport()
        self.ctx.check(self.L.hpp_distance_layer_sharded(self.h, _ptr(occ_band_dev), W, total_rows,
                                                         C.c_float(float(np.float32(resolution))), halo,
                                                         _ptr(band_out_dev), _ptr(full_out_dev), C.byref(rep)))
        return rep

    def close(self):
        if getattr(self, "h", None):
            self.L.hpp_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def plan_batch_multi(maps, map_of_query, variant, starts, goals, traj_cap=1 << 22, traj_out=None):
    """hpp_plan_batch_multi: query q runs on maps[map_of_query[q]] (all maps of one context); one call, no host threads"""
    starts = np.ascontiguousarray(starts, np.float64).reshape(-1, 3)
    goals = np.ascontiguousarray(goals, np.float64).reshape(-1, 3)
    moq = np.ascontiguousarray(map_of_query, np.uint32).reshape(-1)
    assert len(moq) == len(starts) == len(goals) and len(maps) > 0
    handles = (C.c_void_p * len(maps))(*[m.h for m in maps])
    plans = np.zeros(len(starts), PLAN_DTYPE)
    traj = traj_out if traj_out is not None else np.empty((traj_cap, 5))
    used = C.c_size_t(0)
    ctx = maps[0].ctx
    ctx.check(maps[0].L.hpp_plan_batch_multi(C.cast(handles, C.c_void_p), len(maps), moq.ctypes.data, variant,
                                             starts.ctypes.data, goals.ctypes.data, len(starts), plans.ctypes.data,
                                             traj.ctypes.data, len(traj), C.byref(used)))
    return plans, traj[:used.value]


def make_nodes(xyh, cost=0.0, steer=0.0, direction=1):
    xyh = np.asarray(xyh, np.float64).reshape(-1, 3)
    a = np.zeros(len(xyh), NODE_DTYPE)
    a["x"], a["y"], a["heading"] = xyh[:, 0], xyh[:, 1], xyh[:, 2]
    a["cost"], a["steer"], a["direction"] = cost, steer, direction
    return a
