"""ctypes binding of libadmm_b200.so (include/admm_b200.h) for tests and bench.py.

The product's host language is C++ (host/); this module exists so that the parity tests and the benchmark
can drive the C ABI directly.  It contains no numerics of its own: every call below forwards to the
shared library, and the library refuses to run without a CUDA device (no CPU path).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADMMB_LIB") or os.path.join(os.path.dirname(_HERE), "libadmm_b200.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

SOLVER_DIRECT, SOLVER_PCG = 0, 1
STATE_X, STATE_Z, STATE_U, STATE_PROX, STATE_PROX_ITERS = 0, 1, 2, 3, 4

# every symbol include/admm_b200.h declares (tests check that the library exports them all)
EXPORTS = [
    "admmb_create", "admmb_destroy", "admmb_last_error", "admmb_version", "admmb_set_nodes", "admmb_add_tets",
    "admmb_add_tris", "admmb_add_springs", "admmb_add_bends", "admmb_add_static_anchors", "admmb_add_moving_anchors",
    "admmb_add_collision", "admmb_set_gravity", "admmb_add_explicit_subset", "admmb_add_wind", "admmb_set_solver", "admmb_finalize", "admmb_step",
    "admmb_step_dump", "admmb_debug_local_step", "admmb_debug_global_step", "admmb_step_resident", "admmb_upload_xv", "admmb_download_xv", "admmb_update_anchor_targets",
    "admmb_get_anchor_targets", "admmb_set_batch_weights", "admmb_get_batch_weights", "admmb_recompute_weights",
    "admmb_state_size", "admmb_get_state", "admmb_set_state", "admmb_get_info", "admmb_timing_enable",
    "admmb_timing_read", "admmb_last_region_ms", "admmb_dist_unique_id", "admmb_dist_init",
    "admmb_register_host_buffer", "admmb_unregister_host_buffer", "admmb_download_x_f32", "admmb_set_host_threads", "admmb_step_resident_async", "admmb_sync", "admmb_set_deterministic", "admmb_set_check_finite", "admmb_step_async",
    "admmb_probe_fp64", "admmb_debug_fastmath_selftest", "admmb_enable_explicit",
]


class Info(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("n_batches", C.c_int), ("solver", C.c_int), ("n_rows", C.c_long),
                ("nnz_A", C.c_long), ("nnz_L", C.c_long), ("n_supernodes", C.c_int), ("n_levels", C.c_int),
                ("factor_bytes", C.c_long), ("factor_seconds", C.c_double), ("cg_iters_total", C.c_long),
                ("launches_total", C.c_long), ("elapsed_s", C.c_double), ("device_fronts", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (there is no fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.admmb_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.admmb_destroy.argtypes = [vp]
    L.admmb_last_error.argtypes = [vp]
    L.admmb_last_error.restype = C.c_char_p
    L.admmb_version.restype = C.c_char_p
    L.admmb_set_nodes.argtypes = [vp, C.c_int, _dp, _dp]
    L.admmb_add_tets.argtypes = [vp, C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double, C.c_int]
    L.admmb_add_tris.argtypes = [vp, C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double, C.c_int]
    L.admmb_add_springs.argtypes = [vp, C.c_int, _ip, _dp]
    L.admmb_add_bends.argtypes = [vp, C.c_int, _ip, C.c_double]
    L.admmb_add_static_anchors.argtypes = [vp, C.c_int, _ip, C.c_double]
    L.admmb_add_moving_anchors.argtypes = [vp, C.c_int, _ip, _dp, C.c_double]
    L.admmb_add_collision.argtypes = [vp, C.c_int, _ip, _dp, C.c_double]
    L.admmb_set_gravity.argtypes = [vp, C.c_int, _dp]
    L.admmb_set_host_threads.argtypes = [C.c_int]
    L.admmb_set_deterministic.argtypes = [vp, C.c_int]
    L.admmb_set_check_finite.argtypes = [vp, C.c_int]
    L.admmb_register_host_buffer.argtypes = [vp, vp, C.c_long]
    L.admmb_download_x_f32.argtypes = [vp, np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")]
    L.admmb_unregister_host_buffer.argtypes = [vp, vp]
    L.admmb_add_explicit_subset.argtypes = [vp, C.c_int, _ip, _dp]
    L.admmb_add_wind.argtypes = [vp, C.c_int, _ip, _dp]
    L.admmb_set_solver.argtypes = [vp, C.c_int, C.c_double, C.c_int]
    L.admmb_finalize.argtypes = [vp, C.c_double]
    L.admmb_step.argtypes = [vp, C.c_int, _dp, _dp]
    L.admmb_step_dump.argtypes = [vp, C.c_int, _dp, _dp, vp, vp, vp]
    L.admmb_debug_local_step.argtypes = [vp, _dp]
    L.admmb_debug_global_step.argtypes = [vp, _dp]
    L.admmb_step_resident.argtypes = [vp, C.c_int, C.c_int]
    L.admmb_step_resident_async.argtypes = [vp, C.c_int, C.c_int]
    L.admmb_sync.argtypes = [vp]
    L.admmb_step_async.argtypes = [vp, C.c_int, _dp, _dp]
    L.admmb_upload_xv.argtypes = [vp, vp, vp]
    L.admmb_download_xv.argtypes = [vp, vp, vp]
    L.admmb_update_anchor_targets.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.admmb_get_anchor_targets.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.admmb_set_batch_weights.argtypes = [vp, C.c_int, _dp]
    L.admmb_get_batch_weights.argtypes = [vp, C.c_int, _dp]
    L.admmb_recompute_weights.argtypes = [vp]
    L.admmb_state_size.argtypes = [vp, C.c_int]
    L.admmb_state_size.restype = C.c_long
    L.admmb_get_state.argtypes = [vp, C.c_int, _dp]
    L.admmb_set_state.argtypes = [vp, C.c_int, _dp]
    L.admmb_get_info.argtypes = [vp, C.POINTER(Info)]
    L.admmb_timing_enable.argtypes = [vp, C.c_int]
    L.admmb_timing_read.argtypes = [vp, _dp, C.POINTER(C.c_long), C.c_int]
    L.admmb_last_region_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.admmb_dist_unique_id.argtypes = [C.c_char_p]
    L.admmb_dist_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.admmb_enable_explicit.argtypes = [vp, C.c_int, C.c_int]
    L.admmb_probe_fp64.argtypes = [C.c_int, _dp]
    L.admmb_debug_fastmath_selftest.argtypes = [C.c_int, C.c_ulonglong, C.c_long, np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")]
    _lib = L
    return L


def probe_fp64(device=0):
    """Measured FP64 issue rates on `device`: dict with DFMA / DADD / DMUL 1e12 thread-instructions/s, the dependent-DFMA
    latency in cycles, SM count and max SM clock (admmb_probe_fp64)."""
    out = np.zeros(6)
    rc = lib().admmb_probe_fp64(int(device), out)
    if rc != 0:
        raise RuntimeError(f"admmb_probe_fp64 failed ({rc})")
    return {"dfma_tinst_s": out[0], "dadd_tinst_s": out[1], "dmul_tinst_s": out[2], "dfma_dependent_cycles": out[3],
            "sms": int(out[4]), "sm_max_mhz": out[5]}


def fastmath_selftest(samples=1 << 26, seed=1, device=0):
    """(mismatches[4], fallbacks[4]) of the exact fast division / reciprocal / sqrt paths against the operators."""
    out = np.zeros(8, dtype=np.uint64)
    rc = lib().admmb_debug_fastmath_selftest(int(device), int(seed), int(samples), out)
    if rc != 0:
        raise RuntimeError(f"admmb_debug_fastmath_selftest failed ({rc})")
    return out[:4].astype(np.int64), out[4:].astype(np.int64)


_flop_lib = None


def flopcount_hyper_tets(kind, x_rest, idx, mu, lam, maxit, dt, x_cur, u, state):
    """Algorithmic FP64 flops of ONE ADMM iteration's local step over the given hyperelastic tets (csrc/flopcount.cpp: the
    kernels' per-force body compiled for the host with counters; add / mul / div / sqrt / log = 1 flop).  idx [count][4],
    u [count][9], state [count][4] entering the iteration.  Returns (flops, objective evaluations, L-BFGS iterations)."""
    global _flop_lib
    if _flop_lib is None:
        path = os.path.join(os.path.dirname(_HERE), "libadmm_b200_flopcount.so")
        _flop_lib = C.CDLL(path)
        _flop_lib.admmb_flopcount_hyper_tets.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _ip, C.c_double, C.c_double, C.c_int, C.c_double,
                                                        _dp, _dp, _dp, _dp]
    idx = _i32(idx).reshape(-1, 4)
    out = np.zeros(3)
    x_rest = _f64(x_rest).reshape(-1)
    rc = _flop_lib.admmb_flopcount_hyper_tets(int(kind), x_rest.size // 3, x_rest, idx.shape[0], idx, float(mu), float(lam), int(maxit), float(dt),
                                              _f64(x_cur).reshape(-1), _f64(u).reshape(-1), _f64(state).reshape(-1), out)
    if rc != 0:
        raise RuntimeError("admmb_flopcount_hyper_tets failed")
    return float(out[0]), float(out[1]), float(out[2])


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class AdmmError(RuntimeError):
    pass


class System:
    """Mirror of admm::System (A/src/system/System.hpp:29-99) on top of the C ABI.

    Public members follow the reference: m_x, m_v (numpy, 3n), settings via dt / admm_iters, step(),
    recompute_weights().  Built from a scene dictionary (scenes.py).
    """

    def __init__(self, scene, device=0, solver=SOLVER_DIRECT, cg_tol=1e-12, cg_max_iters=20000, iters=None, dist=None, host_explicit=False, pin_host=False, deterministic=None):
        """dist = (rank, world, id128 bytes) partitions the mesh over `world` processes (PCG only)."""
        L = lib()
        self.L = L
        h = C.c_void_p()
        rc = L.admmb_create(int(device), C.byref(h))
        if rc != 0:
            raise AdmmError(f"admmb_create failed ({rc}): {L.admmb_last_error(None).decode()}")
        self.h = h
        self.scene = scene
        self.dt = float(scene["dt"])
        self.admm_iters = int(iters if iters is not None else scene["iters"])
        x = _f64(scene["x"]).reshape(-1)
        self.n3 = x.size
        self.m_x = x.copy()
        self.m_v = np.zeros_like(x)
        self._ck(L.admmb_set_nodes(h, x.size // 3, x, _f64(np.repeat(_f64(scene["m"]).reshape(-1), 3))))
        self.batch_ids = []
        for b in scene["batches"]:
            t = b["type"]
            if t == "tets":
                idx = _i32(b["idx"])
                bid = L.admmb_add_tets(h, int(b["kind"]), idx.shape[0], idx, float(b.get("p0", 0)), float(b.get("p1", 0)),
                                       float(b.get("p2", 0)), int(b.get("maxit", 10)))
            elif t == "tris":
                idx = _i32(b["idx"])
                bid = L.admmb_add_tris(h, int(b["kind"]), idx.shape[0], idx, float(b["stiffness"]),
                                       float(b.get("lmin", 0.0)), float(b.get("lmax", 9999999.0)), int(b.get("flag", 1)))
            elif t == "springs":
                idx = _i32(b["idx"])
                bid = L.admmb_add_springs(h, idx.shape[0], idx, _f64(np.broadcast_to(b["stiffness"], (idx.shape[0],))))
            elif t == "bends":
                idx = _i32(b["idx"])
                bid = L.admmb_add_bends(h, idx.shape[0], idx, float(b["stiffness"]))
            elif t == "static_anchors":
                idx = _i32(b["idx"])
                bid = L.admmb_add_static_anchors(h, idx.size, idx, float(b.get("weight", -1.0)))
            elif t == "moving_anchors":
                idx = _i32(b["idx"])
                bid = L.admmb_add_moving_anchors(h, idx.size, idx, _f64(b["pos"]), float(b.get("weight", -1.0)))
            elif t == "collision":
                kinds = _i32(b["kinds"])
                bid = L.admmb_add_collision(h, kinds.size, kinds, _f64(b["params"]), float(b.get("weight", 32.0)))
            else:
                raise ValueError(t)
            if bid < 0:
                self._ck(bid)
            self.batch_ids.append(bid)
        # Explicit forces run in list order (System.cpp:37-39), on the device: gravity (all nodes or a subset) and
        # wind.  host_explicit=True keeps them on the host instead (what a user-defined ExplicitForce subclass gets:
        # the caller applies them to m_v before step(), see apply_host_explicit) -- used by the tests as a cross-check.
        self.gravity_ids = []
        self.host_explicit = []
        ex = scene.get("explicit", [])
        if host_explicit:
            self.host_explicit = list(ex)
        else:
            for e in ex:
                if e["type"] == "gravity" and e.get("indices") is not None and len(e["indices"]):
                    ii = _i32(e["indices"])
                    gid = L.admmb_add_explicit_subset(h, ii.size, ii, _f64(e["dir"]))
                elif e["type"] == "gravity":
                    gid = L.admmb_set_gravity(h, -1, _f64(e["dir"]))
                elif e["type"] == "wind":
                    tt = _i32(e["tris"])
                    gid = L.admmb_add_wind(h, tt.size // 3, tt, _f64(e["dir"]))
                else:
                    raise ValueError(e["type"])
                if gid < 0:
                    self._ck(gid)
                self.gravity_ids.append(gid)
        self._ck(L.admmb_set_solver(h, int(solver), float(cg_tol), int(cg_max_iters)))
        if deterministic is not None:
            self._ck(L.admmb_set_deterministic(h, 1 if deterministic else 0))
        if dist is not None:
            self._ck(L.admmb_dist_init(h, int(dist[0]), int(dist[1]), dist[2]))
        self._ck(L.admmb_finalize(h, self.dt))
        # pin_host: page-lock m_x / m_v so that step() transfers them directly (no staging copy); the arrays are kept
        # alive until close() even if the caller rebinds self.m_x
        self._pinned = []
        if pin_host:
            for a in (self.m_x, self.m_v):
                self._ck(L.admmb_register_host_buffer(h, a.ctypes.data_as(C.c_void_p), a.nbytes))
                self._pinned.append(a)

    def _ck(self, rc):
        if rc < 0:
            raise AdmmError(f"libadmm_b200 error {rc}: {self.L.admmb_last_error(self.h).decode()}")
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.L.admmb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply_host_explicit(self):
        """ExplicitForce::project / WindForce::project for forces kept on the host (per frame, before step)."""
        for e in self.host_explicit:
            if e["type"] == "gravity" and e.get("indices") is not None and len(e["indices"]):
                for i in e["indices"]:
                    self.m_v.reshape(-1, 3)[int(i)] += self.dt * _f64(e["dir"])
            elif e["type"] == "gravity":
                self.m_v.reshape(-1, 3)[:] += self.dt * _f64(e["dir"])
            else:
                wind_project(self.m_x, self.m_v, e["tris"], e["dir"], self.dt)

    # --- System::step() ---
    def step(self, iters=None):
        self._ck(self.L.admmb_step(self.h, int(self.admm_iters if iters is None else iters), self.m_x, self.m_v))

    def step_dump(self, iters=None):
        """One step with per-iteration dumps: (x_it [K,3n], z_it [K,R], u_it [K,R]); m_x/m_v updated."""
        K = int(self.admm_iters if iters is None else iters)
        R = self.L.admmb_state_size(self.h, STATE_Z)
        xi = np.zeros((K, self.n3))
        zi = np.zeros((K, max(R, 0)))
        ui = np.zeros((K, max(R, 0)))
        self._ck(self.L.admmb_step_dump(self.h, K, self.m_x, self.m_v, xi.ctypes.data_as(C.c_void_p),
                                        zi.ctypes.data_as(C.c_void_p), ui.ctypes.data_as(C.c_void_p)))
        return xi, zi, ui

    def debug_local_step(self, x):
        self._ck(self.L.admmb_debug_local_step(self.h, _f64(x).reshape(-1)))

    def debug_global_step(self, xbar):
        self._ck(self.L.admmb_debug_global_step(self.h, _f64(xbar).reshape(-1)))

    def step_resident(self, frames=1, iters=None):
        self._ck(self.L.admmb_step_resident(self.h, int(self.admm_iters if iters is None else iters), int(frames)))

    def step_resident_async(self, frames=1, iters=None):
        """Enqueue only (scene ensembles: several Systems in flight on one GPU); pair with sync()."""
        self._ck(self.L.admmb_step_resident_async(self.h, int(self.admm_iters if iters is None else iters), int(frames)))

    def step_async(self, iters=None):
        """System::step() enqueued only: m_x / m_v are read now and written when sync() returns."""
        self._ck(self.L.admmb_step_async(self.h, int(self.admm_iters if iters is None else iters), self.m_x, self.m_v))

    def sync(self):
        self._ck(self.L.admmb_sync(self.h))

    def upload(self):
        self._ck(self.L.admmb_upload_xv(self.h, self.m_x.ctypes.data_as(C.c_void_p), self.m_v.ctypes.data_as(C.c_void_p)))

    def download(self):
        self._ck(self.L.admmb_download_xv(self.h, self.m_x.ctypes.data_as(C.c_void_p), self.m_v.ctypes.data_as(C.c_void_p)))

    def positions_f32(self):
        """Render hand-off: current device positions as float32 [n, 3] (SimContext::update's float sink)."""
        out = np.empty(self.n3, dtype=np.float32)
        self._ck(self.L.admmb_download_x_f32(self.h, out))
        return out.reshape(-1, 3)

    def set_x(self, x):
        self.m_x[:] = _f64(x).reshape(-1)

    # --- state ---
    def state(self, which):
        n = self.L.admmb_state_size(self.h, which)
        out = np.zeros(max(n, 0), dtype=np.float64)
        if n > 0:
            self._ck(self.L.admmb_get_state(self.h, which, out))
        return out

    def set_state(self, which, a):
        self._ck(self.L.admmb_set_state(self.h, which, _f64(a).reshape(-1)))

    x_iter = property(lambda s: s.state(STATE_X))
    z = property(lambda s: s.state(STATE_Z))
    u = property(lambda s: s.state(STATE_U))

    def prox_state(self):
        return self.state(STATE_PROX).reshape(-1, 4)

    def prox_iters(self):
        return self.state(STATE_PROX_ITERS).astype(np.int32)

    # --- runtime changes ---
    def update_anchor_targets(self, batch, first, pos=None, active=None):
        cnt = len(pos) if pos is not None else len(active)
        p = None if pos is None else _f64(pos)
        a = None if active is None else _i32(active)
        self._ck(self.L.admmb_update_anchor_targets(self.h, batch, first, cnt,
                                                    None if p is None else p.ctypes.data_as(C.c_void_p),
                                                    None if a is None else a.ctypes.data_as(C.c_void_p)))

    def get_anchor_targets(self, batch, first, count):
        p = np.zeros((count, 3))
        a = np.zeros(count, dtype=np.int32)
        self._ck(self.L.admmb_get_anchor_targets(self.h, batch, first, count, p.ctypes.data_as(C.c_void_p),
                                                 a.ctypes.data_as(C.c_void_p)))
        return p, a

    def set_batch_weights(self, batch, w):
        self._ck(self.L.admmb_set_batch_weights(self.h, batch, _f64(w)))

    def get_batch_weights(self, batch, count):
        w = np.zeros(count)
        self._ck(self.L.admmb_get_batch_weights(self.h, batch, w))
        return w

    def recompute_weights(self):
        self._ck(self.L.admmb_recompute_weights(self.h))

    def set_gravity(self, gid, d):
        self._ck(self.L.admmb_set_gravity(self.h, gid, _f64(d)))

    # --- measurement ---
    def info(self):
        i = Info()
        self._ck(self.L.admmb_get_info(self.h, C.byref(i)))
        return {f: getattr(i, f) for f, _ in Info._fields_}

    def last_region_ms(self):
        ms = C.c_double(0.0)
        self._ck(self.L.admmb_last_region_ms(self.h, C.byref(ms)))
        return ms.value

    def timing(self, on):
        self._ck(self.L.admmb_timing_enable(self.h, 1 if on else 0))

    def timing_read(self, reset=True):
        ms = np.zeros(4)
        it = C.c_long(0)
        self._ck(self.L.admmb_timing_read(self.h, ms, C.byref(it), 1 if reset else 0))
        return dict(local_ms=ms[0], rhs_ms=ms[1], solve_ms=ms[2], step_ms=ms[3], iters=it.value)


def dist_unique_id():
    """ncclUniqueId (128 bytes) to be created on rank 0 and handed to every rank's System(dist=...)."""
    buf = C.create_string_buffer(128)
    rc = lib().admmb_dist_unique_id(buf)
    if rc != 0:
        raise AdmmError(f"admmb_dist_unique_id failed ({rc})")
    return buf.raw


def wind_project(x, v, tris, direction, dt):
    """WindForce::project (A/src/system/ExplicitForce.cpp:42-98), serial triangle order (OMP_NUM_THREADS=1):
    a per-frame explicit force the caller applies to m_v before step(), like any user ExplicitForce."""
    x = x.reshape(-1, 3)
    v = v.reshape(-1, 3)
    d = np.asarray(direction, dtype=np.float64)
    for t in tris:
        i0, i1, i2 = int(t[0]), int(t[1]), int(t[2])
        curr_v = (v[i0] + v[i1] + v[i2]) / 3.0
        v_r = curr_v - d
        n = np.cross(x[i1] - x[i0], x[i2] - x[i0])
        nn = np.sqrt(n[0] * n[0] + (n[1] * n[1] + n[2] * n[2]))
        normal = n / nn
        area = 0.5 * nn
        v_n = normal[0] * v_r[0] + (normal[1] * v_r[1] + normal[2] * v_r[2])
        force = -1000.0 * area * v_n * abs(v_n) * normal
        force = force * 0.33
        force = force * dt
        v[i0] += force
        v[i1] += force
        v[i2] += force
