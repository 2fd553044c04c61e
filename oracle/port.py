"""ctypes view of oracle/libadmm_oracle.so -- the plain-C restatement of the reference's solver path
(oracle/port/admm_oracle.c).  TEST INFRASTRUCTURE ONLY (same rules as oracle/ref.py)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadmm_oracle.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = _vp
        L.oracle_create.argtypes = [C.c_int, _dp, _dp, C.c_double]
        L.oracle_destroy.argtypes = [_vp]
        L.oracle_add_tets.argtypes = [_vp, C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double, C.c_int]
        L.oracle_add_tris.argtypes = [_vp, C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_double, C.c_int]
        L.oracle_add_springs.argtypes = [_vp, C.c_int, _ip, _dp]
        L.oracle_add_bends.argtypes = [_vp, C.c_int, _ip, C.c_double]
        L.oracle_add_anchors.argtypes = [_vp, C.c_int, C.c_int, _ip, _vp, C.c_double]
        L.oracle_set_anchor.argtypes = [_vp, C.c_int, _vp, C.c_int]
        L.oracle_get_anchor.argtypes = [_vp, C.c_int, _dp]
        L.oracle_set_weight.argtypes = [_vp, C.c_int, C.c_double]
        L.oracle_add_collision.argtypes = [_vp, C.c_int, _ip, _dp, C.c_double]
        L.oracle_add_gravity.argtypes = [_vp, _dp]
        L.oracle_add_wind.argtypes = [_vp, C.c_int, _ip, _dp]
        L.oracle_initialize.argtypes = [_vp]
        L.oracle_rows.restype = C.c_long
        L.oracle_rows.argtypes = [_vp]
        L.oracle_num_hyper.argtypes = [_vp]
        L.oracle_get_prox.argtypes = [_vp, _vp, _vp]
        L.oracle_step.argtypes = [_vp, C.c_int, _dp, _dp, _vp, _vp, _vp, _vp]
        L.oracle_local_step.argtypes = [_vp, _dp, _dp, _vp, _dp, _dp, _vp]
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class PortAdapter:
    """Same scripting interface as scenarios.RefAdapter / DevAdapter, on the C restatement."""

    def __init__(self, scene, iters=None):
        L = lib()
        self.L, self.scene = L, scene
        self.iters = int(iters if iters is not None else scene["iters"])
        x = _f64(scene["x"]).reshape(-1)
        self.x = x.copy()
        self.v = np.zeros_like(x)
        self.h = L.oracle_create(x.size // 3, x, _f64(scene["m"]).reshape(-1), float(scene["dt"]))
        self.first = {}
        for bi, b in enumerate(scene["batches"]):
            t = b["type"]
            idx = None if t == "collision" else _i32(b["idx"])
            if t == "tets":
                L.oracle_add_tets(self.h, int(b["kind"]), idx.shape[0], idx, float(b.get("p0", 0)), float(b.get("p1", 0)),
                                  float(b.get("p2", 0)), int(b.get("maxit", 10)))
            elif t == "tris":
                L.oracle_add_tris(self.h, int(b["kind"]), idx.shape[0], idx, float(b["stiffness"]), float(b.get("lmin", 0.0)),
                                  float(b.get("lmax", 9999999.0)), int(b.get("flag", 1)))
            elif t == "springs":
                L.oracle_add_springs(self.h, idx.shape[0], idx, _f64(np.broadcast_to(b["stiffness"], (idx.shape[0],))))
            elif t == "bends":
                L.oracle_add_bends(self.h, idx.shape[0], idx, float(b["stiffness"]))
            elif t == "static_anchors":
                self.first[bi] = L.oracle_add_anchors(self.h, 0, idx.size, idx, None, float(b.get("weight", -1.0)))
            elif t == "moving_anchors":
                pos = _f64(b["pos"])
                self.first[bi] = L.oracle_add_anchors(self.h, 1, idx.size, idx, pos.ctypes.data_as(_vp), float(b.get("weight", -1.0)))
            elif t == "collision":
                assert bi == len(scene["batches"]) - 1, "the port applies the collision force last"
                L.oracle_add_collision(self.h, len(b["kinds"]), _i32(b["kinds"]), _f64(b["params"]), float(b.get("weight", 32.0)))
        for e in scene.get("explicit", []):
            if e["type"] == "gravity":
                L.oracle_add_gravity(self.h, _f64(e["dir"]))
            else:
                tr = _i32(e["tris"])
                L.oracle_add_wind(self.h, tr.shape[0], tr, _f64(e["dir"]))
        if L.oracle_initialize(self.h) != 0:
            raise RuntimeError("oracle port: matrix not positive definite")
        self.rows = L.oracle_rows(self.h)
        self.nh = L.oracle_num_hyper(self.h)

    def set_x(self, x):
        self.x[:] = _f64(x).reshape(-1)

    def set_control_points(self, batch, pos=None, active=None):
        cnt = len(self.scene["batches"][batch]["idx"])
        for i in range(cnt):
            p = None if pos is None else _f64(pos[i]).ctypes.data_as(_vp)
            self.L.oracle_set_anchor(self.h, self.first[batch] + i, p, -1 if active is None else int(active[i]))

    def get_control_points(self, batch):
        cnt = len(self.scene["batches"][batch]["idx"])
        out = np.zeros((cnt, 3))
        for i in range(cnt):
            self.L.oracle_get_anchor(self.h, self.first[batch] + i, out[i])
        return out

    def set_anchor_weights(self, batch, w):
        cnt = len(self.scene["batches"][batch]["idx"])
        for i, wi in enumerate(np.broadcast_to(w, (cnt,))):
            self.L.oracle_set_weight(self.h, self.first[batch] + i, float(wi))

    def recompute_weights(self):
        if self.L.oracle_initialize(self.h) != 0:
            raise RuntimeError("oracle port: refactorisation failed")

    def step_dump(self):
        K = self.iters
        xi = np.zeros((K, self.x.size))
        zi = np.zeros((K, self.rows))
        ui = np.zeros((K, self.rows))
        pi = np.zeros((K, self.nh, 4))
        self.L.oracle_step(self.h, K, self.x, self.v, xi.ctypes.data_as(_vp), zi.ctypes.data_as(_vp), ui.ctypes.data_as(_vp),
                           pi.ctypes.data_as(_vp) if self.nh else None)
        self.last_prox_it = pi
        return xi, zi, ui, self.x.copy(), self.v.copy()

    def step(self):
        self.L.oracle_step(self.h, self.iters, self.x, self.v, None, None, None, None)
        return self.x.copy(), self.v.copy()

    def local_step(self, x, u, prox=None):
        """Teacher-forced half iteration: (z, u, optimiser state) of the local step on curr_x = x from the given u / state."""
        z, uo = np.zeros(self.rows), np.zeros(self.rows)
        po = np.zeros((self.nh, 4))
        pin = None if (prox is None or not self.nh) else _f64(prox).ctypes.data_as(_vp)
        self.L.oracle_local_step(self.h, _f64(x).reshape(-1), _f64(u).reshape(-1), pin, z, uo, po.ctypes.data_as(_vp) if self.nh else None)
        return z, uo, po

    def prox_state(self):
        out = np.zeros((self.nh, 4))
        if self.nh:
            self.L.oracle_get_prox(self.h, out.ctypes.data_as(_vp), None)
        return out

    def prox_iters(self):
        out = np.zeros(self.nh, dtype=np.int32)
        if self.nh:
            self.L.oracle_get_prox(self.h, None, out.ctypes.data_as(_vp))
        return out

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None
