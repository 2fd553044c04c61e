"""ctypes view of oracle/_ref/libadmm_ref.so -- the UNMODIFIED reference solver.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never
by the product path.  The library is built by oracle/Makefile from the sources under
/root/reference (where they lie); see oracle/ref_shim.cpp for what each call maps to.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libadmm_ref.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_double, C.c_int, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_add_nodes.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.ref_add_tets.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double, C.c_int]
        L.ref_add_tris.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double, C.c_int]
        L.ref_add_springs.argtypes = [C.c_void_p, C.c_int, _ip, _dp]
        L.ref_add_bends.argtypes = [C.c_void_p, C.c_int, _ip, C.c_double]
        L.ref_add_static_anchors.argtypes = [C.c_void_p, C.c_int, _ip, C.c_double]
        L.ref_add_moving_anchors.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_double]
        L.ref_set_control_point.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_get_control_point.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_set_moving_anchor_weight.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ref_recompute_weights.argtypes = [C.c_void_p]
        L.ref_add_collision.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_double]
        L.ref_add_explicit.argtypes = [C.c_void_p, _dp]
        L.ref_add_wind.argtypes = [C.c_void_p, C.c_int, _ip, _dp]
        L.ref_set_explicit_direction.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_initialize.argtypes = [C.c_void_p, C.c_int]
        for f in ("ref_num_dof", "ref_live_rows", "ref_D_rows", "ref_D_nnz", "ref_L_nnz"):
            getattr(L, f).restype = C.c_long
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_elapsed.restype = C.c_double
        L.ref_elapsed.argtypes = [C.c_void_p]
        L.ref_set_iters.argtypes = [C.c_void_p, C.c_int]
        for f in ("ref_get_x", "ref_set_x", "ref_get_v", "ref_set_v", "ref_get_z", "ref_get_u", "ref_set_u",
                  "ref_get_force_weights"):
            getattr(L, f).argtypes = [C.c_void_p, _dp]
        L.ref_get_prox_state.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_set_prox_state.argtypes = [C.c_void_p, _dp]
        L.ref_get_prox_iters.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_step.argtypes = [C.c_void_p]
        L.ref_step_dump.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_void_p]
        L.ref_step_timed.restype = C.c_double
        L.ref_step_timed.argtypes = [C.c_void_p, C.c_int]
        L.ref_set_omp_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class RefSystem:
    """admm::System of the reference, driven through the shim.

    Built from the same scene dictionary (see scenes.py) that the CUDA path consumes.
    """

    def __init__(self, scene, probe=False, verbose=0, iters=None):
        L = lib()
        self.L = L
        self.scene = scene
        self.iters = int(iters if iters is not None else scene["iters"])
        self.h = L.ref_create(float(scene["dt"]), self.iters, verbose)
        x = _f64(scene["x"]).reshape(-1)
        m = _f64(np.repeat(_f64(scene["m"]).reshape(-1), 3))
        L.ref_add_nodes(self.h, x, m, x.size)
        self.n3 = x.size
        self.cp_first = []
        for b in scene["batches"]:
            t = b["type"]
            if t == "tets":
                idx = _i32(b["idx"])
                L.ref_add_tets(self.h, int(b["kind"]), idx.shape[0], idx, float(b.get("p0", 0)),
                               float(b.get("p1", 0)), float(b.get("p2", 0)), int(b.get("maxit", 10)))
            elif t == "tris":
                idx = _i32(b["idx"])
                L.ref_add_tris(self.h, int(b["kind"]), idx.shape[0], idx, float(b["stiffness"]),
                               float(b.get("lmin", 0.0)), float(b.get("lmax", 9999999.0)), int(b.get("flag", 1)))
            elif t == "springs":
                idx = _i32(b["idx"])
                k = _f64(np.broadcast_to(b["stiffness"], (idx.shape[0],)))
                L.ref_add_springs(self.h, idx.shape[0], idx, k)
            elif t == "bends":
                idx = _i32(b["idx"])
                L.ref_add_bends(self.h, idx.shape[0], idx, float(b["stiffness"]))
            elif t == "static_anchors":
                idx = _i32(b["idx"])
                L.ref_add_static_anchors(self.h, idx.size, idx, float(b.get("weight", -1.0)))
            elif t == "moving_anchors":
                idx = _i32(b["idx"])
                pos = _f64(b["pos"])
                self.cp_first.append(L.ref_add_moving_anchors(self.h, idx.size, idx, pos, float(b.get("weight", -1.0))))
            elif t == "collision":
                kinds = _i32(b["kinds"])
                L.ref_add_collision(self.h, kinds.size, kinds, _f64(b["params"]), float(b.get("weight", 32.0)))
            else:
                raise ValueError(t)
        self.explicit = []
        for e in scene.get("explicit", []):
            if e["type"] == "gravity":
                L.ref_add_explicit(self.h, _f64(e["dir"]))
                self.explicit.append(len(self.explicit))
            elif e["type"] == "wind":
                tris = _i32(e["tris"])
                self.explicit.append(L.ref_add_wind(self.h, tris.shape[0], tris, _f64(e["dir"])))
        self.probe = bool(probe)
        if L.ref_initialize(self.h, 1 if probe else 0) != 0:
            raise RuntimeError("reference initialize() failed")
        self.rows = L.ref_live_rows(self.h)

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # state
    def _get(self, fn, n):
        a = np.empty(n, dtype=np.float64)
        getattr(self.L, fn)(self.h, a)
        return a

    x = property(lambda s: s._get("ref_get_x", s.n3))
    v = property(lambda s: s._get("ref_get_v", s.n3))
    z = property(lambda s: s._get("ref_get_z", s.rows))
    u = property(lambda s: s._get("ref_get_u", s.rows))

    def set_x(self, x):
        self.L.ref_set_x(self.h, _f64(x).reshape(-1))

    def set_v(self, v):
        self.L.ref_set_v(self.h, _f64(v).reshape(-1))

    def set_control_point(self, cp, pos=None, active=True):
        p = None if pos is None else _f64(pos).ctypes.data_as(C.c_void_p)
        self.L.ref_set_control_point(self.h, int(cp), p, 1 if active else 0)

    def get_control_point(self, cp):
        p = np.empty(3)
        a = self.L.ref_get_control_point(self.h, int(cp), p)
        return p, bool(a)

    def prox_state(self):
        n = self.L.ref_get_prox_state(self.h, None)
        out = np.empty((n, 4), dtype=np.float64)
        if n:
            self.L.ref_get_prox_state(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def set_prox_state(self, st):
        self.L.ref_set_prox_state(self.h, _f64(st).reshape(-1))

    def set_u(self, u):
        self.L.ref_set_u(self.h, _f64(u).reshape(-1))

    def prox_iters(self):
        n = self.L.ref_get_prox_iters(self.h, None)
        out = np.empty(n, dtype=np.int32)
        if n:
            self.L.ref_get_prox_iters(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def step(self):
        if self.L.ref_step(self.h) != 0:
            raise RuntimeError("reference step() failed")

    def step_dump(self):
        """One unmodified step(); returns (x_it [K,3n], z_it [K,R], u_it [K,R], x_final)."""
        assert self.probe
        K = self.iters
        xi = np.zeros((K, self.n3))
        zi = np.zeros((K, self.rows))
        ui = np.zeros((K, self.rows))
        nh = self.L.ref_get_prox_state(self.h, None)
        pi = np.zeros((K, nh, 4))
        if self.L.ref_step_dump(self.h, xi.reshape(-1), zi.reshape(-1), ui.reshape(-1),
                                pi.ctypes.data_as(C.c_void_p) if nh else None) != 0:
            raise RuntimeError("reference step_dump failed")
        self.last_prox_it = pi
        return xi, zi, ui, self.x

    def step_timed(self, frames):
        return self.L.ref_step_timed(self.h, int(frames))

    def stats(self):
        return dict(D_rows=self.L.ref_D_rows(self.h), D_nnz=self.L.ref_D_nnz(self.h), L_nnz=self.L.ref_L_nnz(self.h))
