"""oracle/pyorc.py -- TEST INFRASTRUCTURE: ctypes bindings for the CPU oracle (liborc.so) and, when
present, the compiled reference (oracle/_ref/libref_{rs,db}.so).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CAR = (0.6, 2.75, 1.1, 1.2, 1.9)  # GL/config/reeds_config.yaml:1-9
RES_F32 = float(np.float32(0.2))          # hazard H6: (double)0.2f


class OrcCar(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("max_steer", "wheel_base", "axle_to_front", "axle_to_back", "width")]


class OrcMap(C.Structure):
    _fields_ = [("occ", C.c_void_p), ("W", C.c_int), ("H", C.c_int), ("res", C.c_double),
                ("ox", C.c_double), ("oy", C.c_double), ("distance", C.c_void_p)]


class OrcState(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("x", "y", "heading", "gridx", "gridy")]


class OrcChild(C.Structure):
    _fields_ = [("traj", OrcState * 8), ("n_traj", C.c_int), ("cost", C.c_double), ("steer", C.c_double),
                ("direction", C.c_int), ("valid", C.c_int)]


class OrcQuirks(C.Structure):
    _fields_ = [("h1_fired", C.c_int), ("h2_clamped", C.c_int), ("processed", C.c_int64), ("max_cost", C.c_double)]


class OrcRsPath(C.Structure):
    _fields_ = [("n_seg", C.c_int), ("lengths", C.c_double * 5), ("types", C.c_char * 5), ("L", C.c_double),
                ("n", C.c_int), ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)),
                ("yaw", C.POINTER(C.c_double)), ("dir", C.POINTER(C.c_int))]


class OrcRsPaths(C.Structure):
    _fields_ = [("n", C.c_int), ("p", OrcRsPath * 48)]


class OrcShot(C.Structure):
    _fields_ = [("found", C.c_int), ("cost", C.c_double), ("n", C.c_int), ("traj", C.POINTER(OrcState)),
                ("n_paths", C.c_int), ("chosen", C.c_int)]


class OrcDubins(C.Structure):
    _fields_ = [("length", C.c_double), ("mode", C.c_char * 3), ("t", C.c_double), ("p", C.c_double),
                ("q", C.c_double), ("n", C.c_int), ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)),
                ("yaw", C.POINTER(C.c_double))]


class OrcPlan(C.Structure):
    _fields_ = [("n", C.c_int), ("traj", C.POINTER(OrcState)), ("iterations", C.c_int), ("via_shot", C.c_int),
                ("cost", C.c_double)]


_lib = None


def build(force=False):
    """Compile liborc.so (and oracle/_ref when the reference tree is present)."""
    so = os.path.join(HERE, "liborc.so")
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h", ".cpp", ".hpp", ".sh"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["bash", os.path.join(HERE, "build.sh")], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liborc.so"))
        L.orc_collide_circles.argtypes = [C.POINTER(OrcMap), C.POINTER(OrcCar), C.c_double, C.c_double, C.c_double]
        L.orc_collide_sat.argtypes = [C.POINTER(OrcMap), C.POINTER(OrcCar), C.c_double, C.c_double, C.c_double]
        L.orc_collide_circles_batch.argtypes = [C.POINTER(OrcMap), C.POINTER(OrcCar), C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_collide_sat_batch.argtypes = L.orc_collide_circles_batch.argtypes
        L.orc_collide_circles_batch.restype = L.orc_collide_sat_batch.restype = None
        L.orc_rs_all_paths.argtypes = [C.c_double] * 8 + [C.POINTER(OrcRsPaths)]
        L.orc_dubins_path.argtypes = [C.c_double] * 8 + [C.POINTER(OrcDubins)]
        L.orc_inflate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.orc_edt_layers.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_edt_sq_exact.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_heuristic_rs.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.POINTER(OrcQuirks)]
        L.orc_heuristic_dubins.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]
        L.orc_otsu_threshold.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_ellipse_halfwidths.argtypes = [C.c_int, C.c_void_p]
        L.orc_expand.argtypes = [C.POINTER(OrcMap), C.POINTER(OrcCar), C.c_int, C.c_double, C.c_double,
                                 C.POINTER(OrcState), C.c_double, C.c_double, C.c_int, C.POINTER(OrcChild)]
        L.orc_rs_shot.argtypes = [C.POINTER(OrcMap), C.POINTER(OrcCar), C.POINTER(OrcState), C.c_double,
                                  C.POINTER(OrcState), C.POINTER(OrcShot)]
        L.orc_dubins_shot.argtypes = L.orc_rs_shot.argtypes
        L.orc_run.argtypes = [C.POINTER(OrcMap), C.POINTER(OrcCar), C.c_int, C.c_double, C.c_double] + \
            [C.c_double] * 6 + [C.c_int, C.POINTER(OrcPlan)]
        L.orc_motion_commands.argtypes = [C.POINTER(OrcCar), C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def car_struct(car=DEFAULT_CAR):
    return OrcCar(*car)


def _i8(a):
    a = np.ascontiguousarray(a, dtype=np.int8)
    return a


class Map:
    """Occupancy map + (optionally) its distance layer, as the RS planner holds it."""

    def __init__(self, occ, resolution=0.2, origin=None, with_layers=True):
        self.occ = _i8(occ)
        self.H, self.W = self.occ.shape
        self.res_f = np.float32(resolution)
        self.res = float(self.res_f)
        if origin is None:
            origin = (-self.W * self.res / 2.0, -self.H * self.res / 2.0)
        self.ox, self.oy = float(origin[0]), float(origin[1])
        self.obstacle = self.distance = None
        if with_layers:
            self.obstacle, self.distance = edt_layers(self.occ, self.res_f)
        self.c = OrcMap(self.occ.ctypes.data, self.W, self.H, self.res, self.ox, self.oy,
                        self.distance.ctypes.data if self.distance is not None else None)


def edt_layers(occ, resolution):
    occ = _i8(occ)
    H, W = occ.shape
    obstacle = np.empty(W * H, np.float32)
    distance = np.empty(W * H, np.float32)
    lib().orc_edt_layers(occ.ctypes.data, W, H, C.c_float(float(resolution)), obstacle.ctypes.data, distance.ctypes.data)
    return obstacle, distance


def cv2_layers(occ, resolution, threads=None):
    """The `obstacle` and `distance` layers with the library call the reference itself makes
    (RS/src/Grid_map.cpp:56-74: cv::distanceTransform(DIST_L2, DIST_MASK_PRECISE) of the [x][y] image, cv::flip(-1),
    times the float resolution), through cv2 with IPP off -- bit-identical to edt_layers() (tests/golden/layers_cv2.npz).
    Used by bench.py's reference arm as the CPU layer build; threads = OpenCV worker threads (None: leave as is)."""
    import cv2
    cv2.ipp.setUseIPP(False)
    if threads is not None:
        cv2.setNumThreads(int(threads))
    occ = _i8(occ)
    free = ~((occ > 20) | (occ < 0))
    obstacle = free.astype(np.float32).reshape(-1)
    d = cv2.distanceTransform(np.ascontiguousarray(free.T.astype(np.uint8)), cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
    d = cv2.flip(d, -1)
    distance = np.ascontiguousarray((d * np.float32(resolution)).astype(np.float32).T).reshape(-1)
    return obstacle, distance


def edt_sq_exact(occ):
    occ = _i8(occ)
    H, W = occ.shape
    out = np.empty((H, W), np.int64)
    lib().orc_edt_sq_exact(occ.ctypes.data, W, H, out.ctypes.data)
    return out


def inflate(occ, resolution, radius):
    occ = _i8(occ)
    H, W = occ.shape
    out = np.empty_like(occ)
    lib().orc_inflate(occ.ctypes.data, W, H, float(resolution), float(radius), out.ctypes.data)
    return out


def image_to_occ(img):
    """f1: 8-bit gray image [rows, cols] -> occupancy [H = cols, W = rows] (grid_map_publisher's convention)"""
    img = np.ascontiguousarray(img, np.uint8)
    rows, cols = img.shape
    occ = np.empty((cols, rows), np.int8)
    L = lib()
    L.orc_image_to_occ.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_image_to_occ(img.ctypes.data, rows, cols, occ.ctypes.data)
    return occ


def heuristic_rs(occ, goal, start=(0, 0), quirks=True):
    occ = _i8(occ)
    H, W = occ.shape
    cost = np.empty((H, W), np.float64)
    q = OrcQuirks()
    lib().orc_heuristic_rs(occ.ctypes.data, W, H, goal[0], goal[1], start[0], start[1], int(quirks),
                           cost.ctypes.data, C.byref(q))
    return cost, q


def heuristic_dubins(occ, goal):
    occ = _i8(occ)
    H, W = occ.shape
    cost = np.empty((H, W), np.float64)
    lib().orc_heuristic_dubins(occ.ctypes.data, W, H, goal[0], goal[1], cost.ctypes.data)
    return cost


def collide_circles(m, poses, car=DEFAULT_CAR):
    """isSingleStateCollisionFree per pose (a C loop over the poses, one orc_collide_circles call each)"""
    cs = car_struct(car)
    poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 3)
    out = np.empty(len(poses), np.uint8)
    lib().orc_collide_circles_batch(C.byref(m.c), C.byref(cs), poses.ctypes.data, len(poses), out.ctypes.data)
    return out


def collide_improved(m, poses, car=DEFAULT_CAR):
    """isSingleStateCollisionFreeImproved per pose (the reference's own return convention)"""
    cs = car_struct(car)
    L = lib()
    L.orc_collide_circles_improved.argtypes = [C.POINTER(OrcMap), C.POINTER(OrcCar), C.c_double, C.c_double, C.c_double]
    poses = np.asarray(poses, np.float64).reshape(-1, 3)
    return np.array([L.orc_collide_circles_improved(C.byref(m.c), C.byref(cs), p[0], p[1], p[2]) for p in poses], np.uint8)


def collide_sat(m, poses, car=DEFAULT_CAR):
    cs = car_struct(car)
    poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 3)
    out = np.empty(len(poses), np.uint8)
    lib().orc_collide_sat_batch(C.byref(m.c), C.byref(cs), poses.ctypes.data, len(poses), out.ctypes.data)
    return out


def motion_commands(variant, car=DEFAULT_CAR):
    cs = car_struct(car)
    steer = np.zeros(64)
    d = np.zeros(64)
    n = lib().orc_motion_commands(C.byref(cs), variant, steer.ctypes.data, d.ctypes.data)
    return steer[:n].copy(), d[:n].copy()


def _state(m, x, y, h):
    L = lib()
    L.orc_cell_x.argtypes = [C.POINTER(OrcMap), C.c_double]
    L.orc_cell_y.argtypes = [C.POINTER(OrcMap), C.c_double]
    return OrcState(x, y, h, L.orc_cell_x(C.byref(m.c), x), L.orc_cell_y(C.byref(m.c), y))


def expand(m, variant, pose, cost, steer, direction, car=DEFAULT_CAR, sim_len=4.0, step=1.0):
    """-> list of dicts (one per motion command, in order), each with valid/cost/steer/direction/traj[n,5]"""
    cs = car_struct(car)
    kids = (OrcChild * 64)()
    st = _state(m, *pose)
    n = lib().orc_expand(C.byref(m.c), C.byref(cs), variant, sim_len, step, C.byref(st), cost, steer, direction, kids)
    out = []
    for k in range(n):
        ch = kids[k]
        traj = np.array([[s.x, s.y, s.heading, s.gridx, s.gridy] for s in ch.traj[:ch.n_traj]])
        out.append(dict(valid=ch.valid, cost=ch.cost, steer=ch.steer, direction=ch.direction, traj=traj))
    return out


def rs_all_paths(s, g, maxc, step=1.0):
    ps = OrcRsPaths()
    L = lib()
    L.orc_rs_all_paths(s[0], s[1], s[2], g[0], g[1], g[2], maxc, step, C.byref(ps))
    out = []
    for k in range(ps.n):
        p = ps.p[k]
        n = p.n
        out.append(dict(lengths=np.array(p.lengths[:p.n_seg]), types=bytes(p.types[:p.n_seg]).decode(), L=p.L,
                        x=np.array(p.x[:n]), y=np.array(p.y[:n]), yaw=np.array(p.yaw[:n]),
                        dir=np.array(p.dir[:n])))
    L.orc_rs_paths_free(C.byref(ps))
    return out


def _shot(fn, m, cur, cur_cost, goal, car):
    cs = car_struct(car)
    c = _state(m, *cur)
    g = _state(m, *goal)
    sh = OrcShot()
    fn(C.byref(m.c), C.byref(cs), C.byref(c), cur_cost, C.byref(g), C.byref(sh))
    traj = np.array([[s.x, s.y, s.heading, s.gridx, s.gridy] for s in sh.traj[:sh.n]]) if sh.found else np.zeros((0, 5))
    out = dict(found=sh.found, cost=sh.cost, traj=traj, n_paths=sh.n_paths, chosen=sh.chosen)
    lib().orc_shot_free(C.byref(sh))
    return out


def rs_shot(m, cur, cur_cost, goal, car=DEFAULT_CAR):
    return _shot(lib().orc_rs_shot, m, cur, cur_cost, goal, car)


def dubins_shot(m, cur, cur_cost, goal, car=DEFAULT_CAR):
    return _shot(lib().orc_dubins_shot, m, cur, cur_cost, goal, car)


def dubins_path(s, g, curv, step=0.1):
    d = OrcDubins()
    L = lib()
    L.orc_dubins_path(s[0], s[1], s[2], g[0], g[1], g[2], curv, step, C.byref(d))
    n = d.n
    out = dict(length=d.length, mode=bytes(d.mode).decode().rstrip("\0"), t=d.t, p=d.p, q=d.q,
               x=np.array(d.x[:n]), y=np.array(d.y[:n]), yaw=np.array(d.yaw[:n]))
    L.orc_dubins_free(C.byref(d))
    return out


def run(m, variant, start, goal, car=DEFAULT_CAR, sim_len=4.0, step=1.0, quirks=True):
    cs = car_struct(car)
    p = OrcPlan()
    lib().orc_run(C.byref(m.c), C.byref(cs), variant, sim_len, step, start[0], start[1], start[2],
                  goal[0], goal[1], goal[2], int(quirks), C.byref(p))
    traj = np.array([[s.x, s.y, s.heading, s.gridx, s.gridy] for s in p.traj[:p.n]])
    out = dict(traj=traj, iterations=p.iterations, via_shot=p.via_shot, cost=p.cost)
    lib().orc_plan_free(C.byref(p))
    return out


# ----------------------------------------------------------------------------------------------------
# compiled reference (oracle/_ref), present only where it was built from /root/reference
# ----------------------------------------------------------------------------------------------------
def ref_available():
    return all(os.path.exists(os.path.join(HERE, "_ref", f)) for f in ("libref_rs.so", "libref_db.so"))


class Ref:
    """One reference planner instance (variant 0 = RS, 1 = Dubins) on one map."""

    def __init__(self, variant, occ, resolution=0.2, origin=None, car=DEFAULT_CAR, sim_len=4.0, step=1.0, layers=None):
        """layers = (obstacle, distance) float32 arrays in grid_map layout: RS only, handed to the stand-in Grid_map
        instead of letting its constructor run the oracle's EDT (the reference arm builds them with cv2)."""
        lib()  # libref_rs.so links against liborc.so
        self.variant = variant
        self.p = "ref_rs_" if variant == 0 else "ref_db_"
        self.L = C.CDLL(os.path.join(HERE, "_ref", "libref_rs.so" if variant == 0 else "libref_db.so"))
        self.occ = _i8(occ)
        self.H, self.W = self.occ.shape
        res_f = np.float32(resolution)
        res = float(res_f)
        if origin is None:
            origin = (-self.W * res / 2.0, -self.H * res / 2.0)
        self.ox, self.oy = origin
        self.res_f = res_f
        self.car = (C.c_double * 5)(*car)
        self.sim = int(sim_len)
        f = getattr(self.L, self.p + "create")
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, C.c_void_p, C.c_double, C.c_double]
        if layers is not None and variant == 0:
            ob = np.ascontiguousarray(layers[0], np.float32).reshape(-1)
            di = np.ascontiguousarray(layers[1], np.float32).reshape(-1)
            assert ob.size == di.size == self.W * self.H
            f = self.L.ref_rs_create_layers
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, C.c_void_p, C.c_double,
                          C.c_double, C.c_void_p, C.c_void_p]
            self.ctx = C.c_void_p(f(self.occ.ctypes.data, self.W, self.H, C.c_float(float(res_f)), self.ox, self.oy,
                                    self.car, sim_len, step, ob.ctypes.data, di.ctypes.data))
        else:
            self.ctx = C.c_void_p(f(self.occ.ctypes.data, self.W, self.H, C.c_float(float(res_f)), self.ox, self.oy,
                                    self.car, sim_len, step))

    def __del__(self):
        try:
            f = getattr(self.L, self.p + "destroy")
            f.argtypes = [C.c_void_p]
            f(self.ctx)
        except Exception:
            pass

    def heuristic(self, goal, want=True):
        """holonomicCostsWithObstacles_planning of this planner instance; want=False skips the copy of the result"""
        cost = np.empty((self.H, self.W), np.float64) if want else None
        if self.variant == 0:
            f = self.L.ref_rs_heuristic_ctx
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            f.restype = None
            f(self.ctx, goal[0], goal[1], cost.ctypes.data if want else None)
        else:
            cost = np.empty((self.H, self.W), np.float64)
            f = self.L.ref_db_heuristic
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            f(self.ctx, goal[0], goal[1], cost.ctypes.data)
        return cost

    def collide(self, poses):
        f = getattr(self.L, self.p + "collide_batch")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        f.restype = None
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 3)
        out = np.empty(len(poses), np.uint8)
        f(self.ctx, poses.ctypes.data, len(poses), out.ctypes.data)
        return out

    def expand(self, pose, cost, steer, direction):
        f = getattr(self.L, self.p + "expand")
        f.argtypes = [C.c_void_p] + [C.c_double] * 5 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        traj = np.zeros((64, self.sim, 5))
        cc = np.zeros(64)
        cs = np.zeros(64)
        cd = np.zeros(64, np.int32)
        n = f(self.ctx, pose[0], pose[1], pose[2], cost, steer, direction, traj.ctypes.data, cc.ctypes.data,
              cs.ctypes.data, cd.ctypes.data, self.sim)
        return [dict(cost=cc[k], steer=cs[k], direction=int(cd[k]), traj=traj[k].copy()) for k in range(n)]

    def shot(self, cur, cur_cost, goal, max_states=20000):
        f = getattr(self.L, self.p + "shot")
        f.argtypes = [C.c_void_p] + [C.c_double] * 7 + [C.c_void_p, C.c_void_p, C.c_int]
        traj = np.zeros((max_states, 5))
        oc = C.c_double(0)
        n = f(self.ctx, cur[0], cur[1], cur[2], cur_cost, goal[0], goal[1], goal[2], C.byref(oc), traj.ctypes.data,
              max_states)
        return dict(found=int(n > 0), cost=oc.value, traj=traj[:n].copy())

    def run(self, start, goal, max_states=200000):
        f = getattr(self.L, self.p + "run")
        f.argtypes = [C.c_void_p] + [C.c_double] * 6 + [C.c_void_p, C.c_int]
        traj = np.zeros((max_states, 5))
        n = f(self.ctx, start[0], start[1], start[2], goal[0], goal[1], goal[2], traj.ctypes.data, max_states)
        return traj[:n].copy()


def ref_rs_all_paths(s, g, maxc, step=1.0, max_samples=400000):
    L = C.CDLL(os.path.join(HERE, "_ref", "libref_rs.so"))
    f = L.ref_rs_all_paths
    f.argtypes = [C.c_double] * 8 + [C.c_void_p] * 9 + [C.c_int]
    nseg = np.zeros(64, np.int32)
    lengths = np.zeros((64, 5))
    types = np.zeros((64, 5), np.uint8)
    Ls = np.zeros(64)
    nsamp = np.zeros(64, np.int32)
    xs = np.zeros(max_samples)
    ys = np.zeros(max_samples)
    yaws = np.zeros(max_samples)
    dirs = np.zeros(max_samples, np.int32)
    n = f(s[0], s[1], s[2], g[0], g[1], g[2], maxc, step, nseg.ctypes.data, lengths.ctypes.data, types.ctypes.data,
          Ls.ctypes.data, nsamp.ctypes.data, xs.ctypes.data, ys.ctypes.data, yaws.ctypes.data, dirs.ctypes.data,
          max_samples)
    out = []
    off = 0
    for k in range(n):
        m = nsamp[k]
        out.append(dict(lengths=lengths[k, :nseg[k]].copy(), types=bytes(types[k, :nseg[k]]).decode(), L=Ls[k],
                        x=xs[off:off + m].copy(), y=ys[off:off + m].copy(), yaw=yaws[off:off + m].copy(),
                        dir=dirs[off:off + m].copy()))
        off += m
    return out


def ref_dubins_path(s, g, curv, max_samples=100000):
    L = C.CDLL(os.path.join(HERE, "_ref", "libref_db.so"))
    f = L.ref_db_dubins_path
    f.argtypes = [C.c_double] * 7 + [C.c_void_p] * 5 + [C.c_int]
    length = C.c_double(0)
    mode = C.create_string_buffer(4)
    xs = np.zeros(max_samples)
    ys = np.zeros(max_samples)
    yaws = np.zeros(max_samples)
    n = f(s[0], s[1], s[2], g[0], g[1], g[2], curv, C.byref(length), mode, xs.ctypes.data, ys.ctypes.data,
          yaws.ctypes.data, max_samples)
    return dict(length=length.value, mode=mode.value.decode(), x=xs[:n].copy(), y=ys[:n].copy(), yaw=yaws[:n].copy())
