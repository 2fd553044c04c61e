"""CPU replay of the golden reference dumps through the PRODUCT's own local-step code (csrc/rest_state.cpp +
csrc/local_bodies.h compiled for the host by tests/hostcheck): for every scenario, frame and ADMM iteration the
reference's curr_x / u / optimiser state are fed in and z, u and the L-BFGS state that come out are compared with
the reference's.  This is the device arithmetic without a device; the -m gpu tests repeat it on the B200."""
import ctypes as C
import os

import numpy as np
import pytest

import scenes
from scenarios import build_scenarios
from util import GOLDEN, rel_l2

HC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck", "libpipelinecheck.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_vp = C.c_void_p
TYPE = dict(tets=0, tris=1, springs=2, bends=3, static_anchors=4, moving_anchors=5, collision=6)
ROWS = [9, 6, 3, 9, 3, 3, 3]
SCEN = build_scenarios()


def _lib():
    L = C.CDLL(HC)
    L.hc_batch_local.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                 C.c_double, _vp, _vp, C.c_int, _vp, _vp, C.c_double, _dp, _dp, _vp, _dp, _dp, _vp, _vp, _vp]
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CpuBatch:
    def __init__(self, L, scene, b):
        self.L, self.b = L, b
        self.type = TYPE[b["type"]]
        self.n = scene["x"].shape[0]
        self.x_rest = np.ascontiguousarray(scene["x"], dtype=np.float64).reshape(-1)
        self.dt = float(scene["dt"])
        self.idx = None if self.type == 6 else np.ascontiguousarray(b["idx"], dtype=np.int32)
        self.count = self.n if self.type == 6 else int(np.asarray(b["idx"]).reshape(-1).size // [4, 3, 2, 4, 1, 1, 1][self.type])
        self.rows = ROWS[self.type] * self.count
        self.kind = int(b.get("kind", 0))
        self.hyper = self.type == 0 and self.kind in (1, 2)
        self.state = np.ones((self.count, 4)) if self.hyper else None
        self.pos = np.ascontiguousarray(b["pos"], dtype=np.float64).copy() if self.type == 5 else None
        self.active = np.ones(self.count, dtype=np.int32) if self.type == 5 else None
        self.weight_override = None

    def project(self, x_cur, u_in):
        b, t = self.b, self.type
        z = np.zeros(self.rows)
        u = np.zeros(self.rows)
        stiff = None
        p0 = p1 = p2 = 0.0
        maxit = flag = 0
        aw = -1.0
        kinds = params = None
        ns = 0
        if t == 0:
            p0, p1, p2, maxit = float(b.get("p0", 0)), float(b.get("p1", 0)), float(b.get("p2", 0)), int(b.get("maxit", 10))
        elif t == 1:
            p0, p1, p2, flag = float(b["stiffness"]), float(b.get("lmin", 0.0)), float(b.get("lmax", 9999999.0)), int(b.get("flag", 1))
            if self.kind == 2 and self.state is None:
                self.state = np.ones((self.count, 1))   # FungTriangle: the force's L-BFGS init_hess
        elif t == 2:
            stiff = np.ascontiguousarray(np.broadcast_to(b["stiffness"], (self.count,)), dtype=np.float64)
        elif t == 3:
            p0 = float(b["stiffness"])
        elif t in (4, 5):
            aw = float(b.get("weight", -1.0)) if self.weight_override is None else self.weight_override
        else:
            aw = float(b.get("weight", 32.0))
            kinds = np.ascontiguousarray(b["kinds"], dtype=np.int32)
            params = np.ascontiguousarray(b["params"], dtype=np.float64)
            ns = kinds.size
        pos_out = np.zeros((self.count, 3)) if t == 5 else None
        rc = self.L.hc_batch_local(t, self.kind, self.n, self.x_rest, self.count, _ptr(self.idx), _ptr(stiff), p0, p1, p2, maxit, flag, aw,
                                   _ptr(self.pos), _ptr(self.active), ns, _ptr(kinds), _ptr(params), self.dt,
                                   np.ascontiguousarray(x_cur, dtype=np.float64).reshape(-1), np.ascontiguousarray(u_in),
                                   _ptr(self.state), z, u, None, None, _ptr(pos_out))
        assert rc == 0
        if t == 5:
            self.pos = pos_out
        return z, u


class CpuAdapter:
    """Minimal event sink so that scenario events (control points, weights) can be replayed on the CPU batches."""

    def __init__(self, batches):
        self.batches = batches

    def set_control_points(self, batch, pos=None, active=None):
        cb = self.batches[batch]
        if pos is not None:
            cb.pos = np.ascontiguousarray(pos, dtype=np.float64).copy()
        if active is not None:
            cb.active = np.ascontiguousarray(active, dtype=np.int32).copy()

    def set_anchor_weights(self, batch, w):
        pass  # weights only enter the right-hand side / system matrix, not the local step of an anchor

    def recompute_weights(self):
        pass


@pytest.mark.parametrize("name", list(SCEN))
def test_local_step_replay_on_cpu(name):
    if not os.path.exists(HC):
        pytest.skip("tests/hostcheck not built (run __graft_entry__.build())")
    L = _lib()
    gold = np.load(os.path.join(GOLDEN, f"{name}.ref.npz"))
    scene = scenes.load_scene(os.path.join(GOLDEN, f"{name}.scene.npz"))
    batches = [CpuBatch(L, scene, b) for b in scene["batches"]]
    offs = np.concatenate([[0], np.cumsum([cb.rows for cb in batches])]).astype(int)
    F, K = gold["x_it"].shape[:2]
    assert offs[-1] == gold["z_it"].shape[2]
    ev = SCEN[name].get("events")
    ad = CpuAdapter(batches)
    u_prev = np.zeros(offs[-1])
    hyper = [cb for cb in batches if cb.hyper]
    n_exact = n_total = 0
    worst = 0.0
    for f in range(F):
        if ev is not None:
            ev(f, ad)
        for k in range(K):
            if hyper and (f or k):
                pk = gold["prox_it"][f, k - 1] if k else gold["prox_it"][f - 1, K - 1]
                o = 0
                for cb in hyper:
                    cb.state = np.ascontiguousarray(pk[o:o + cb.count]).copy()
                    o += cb.count
            for i, cb in enumerate(batches):
                z, u = cb.project(gold["x_it"][f, k], u_prev[offs[i]:offs[i + 1]].copy())
                gz, gu = gold["z_it"][f, k, offs[i]:offs[i + 1]], gold["u_it"][f, k, offs[i]:offs[i + 1]]
                exact = np.array_equal(z, gz) and np.array_equal(u, gu)
                n_total += 1
                n_exact += int(exact)
                worst = max(worst, float(np.abs(z - gz).max()), float(np.abs(u - gu).max()))
                # every force class -- triangles included, through the restated Eigen 3x2 JacobiSVD -- is bit-exact
                assert exact, f"{name}: batch {i} ({cb.b['type']}) frame {f} it {k} not bit-exact: z {rel_l2(z, gz):.2e} u {rel_l2(u, gu):.2e}"
            if hyper:
                got = np.concatenate([cb.state for cb in hyper])
                assert np.array_equal(got, gold["prox_it"][f, k]), f"{name}: L-BFGS state differs at frame {f} it {k}"
            u_prev = gold["u_it"][f, k]
    print(f"{name}: {n_exact}/{n_total} batch projections bit-exact, worst abs difference {worst:.2e}")
    assert worst == 0.0 and n_exact == n_total
