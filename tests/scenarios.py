"""Parity scenarios: a scene dictionary, a frame count and optional per-frame events (control point motion,
weight changes), plus two adapters that run a scenario on the reference (oracle/_ref) and on the CUDA path
through the C ABI with the same script.  Used by tests/golden/make_golden.py and the -m gpu parity tests."""
import numpy as np

import scenes
from scenes import (SHAPE_CYLINDER, SHAPE_FLOOR, SHAPE_SPHERE, TET_ARAP, TET_NH, TET_STVK, TET_VOLUME, TRI_AREA,
                    TRI_STRAIN)


# ---- adapters ---------------------------------------------------------------------------------------------
class RefAdapter:
    """admm::System of the unmodified reference (oracle/ref.py)."""

    def __init__(self, scene, iters=None):
        from oracle.ref import RefSystem
        self.sim = RefSystem(scene, probe=True, iters=iters)
        self.scene = scene
        self.mov = [i for i, b in enumerate(scene["batches"]) if b["type"] == "moving_anchors"]

    def set_x(self, x):
        self.sim.set_x(x)

    def set_control_points(self, batch, pos=None, active=None):
        first = self.sim.cp_first[self.mov.index(batch)]
        cnt = len(self.scene["batches"][batch]["idx"])
        for i in range(cnt):
            _, a = self.sim.get_control_point(first + i)
            self.sim.set_control_point(first + i, None if pos is None else pos[i], a if active is None else bool(active[i]))

    def get_control_points(self, batch):
        first = self.sim.cp_first[self.mov.index(batch)]
        cnt = len(self.scene["batches"][batch]["idx"])
        return np.array([self.sim.get_control_point(first + i)[0] for i in range(cnt)])

    def set_anchor_weights(self, batch, w):
        first = self.sim.cp_first[self.mov.index(batch)]
        for i, wi in enumerate(np.broadcast_to(w, (len(self.scene["batches"][batch]["idx"]),))):
            self.sim.L.ref_set_moving_anchor_weight(self.sim.h, first + i, float(wi))

    def recompute_weights(self):
        self.sim.L.ref_recompute_weights(self.sim.h)

    def step_dump(self):
        xi, zi, ui, xf = self.sim.step_dump()
        self.last_prox_it = self.sim.last_prox_it
        return xi, zi, ui, xf, self.sim.v

    def step(self):
        self.sim.step()
        return self.sim.x, self.sim.v

    def prox_state(self):
        return self.sim.prox_state()

    def prox_iters(self):
        return self.sim.prox_iters()

    def close(self):
        self.sim.close()


class DevAdapter:
    """The CUDA path through libadmm_b200.so (admm_b200.py)."""

    def __init__(self, scene, iters=None, **kw):
        import admm_b200
        self.sim = admm_b200.System(scene, iters=iters, **kw)
        self.scene = scene

    def set_x(self, x):
        self.sim.set_x(x)

    def set_control_points(self, batch, pos=None, active=None):
        self.sim.update_anchor_targets(self.sim.batch_ids[batch], 0, pos=pos, active=active)

    def get_control_points(self, batch):
        return self.sim.get_anchor_targets(self.sim.batch_ids[batch], 0, len(self.scene["batches"][batch]["idx"]))[0]

    def set_anchor_weights(self, batch, w):
        cnt = len(self.scene["batches"][batch]["idx"])
        self.sim.set_batch_weights(self.sim.batch_ids[batch], np.broadcast_to(np.float64(w), (cnt,)).copy())

    def recompute_weights(self):
        self.sim.recompute_weights()

    def step_dump(self):
        self.sim.apply_host_explicit()
        xi, zi, ui = self.sim.step_dump()
        return xi, zi, ui, self.sim.m_x.copy(), self.sim.m_v.copy()

    def step(self):
        self.sim.apply_host_explicit()
        self.sim.step()
        return self.sim.m_x.copy(), self.sim.m_v.copy()

    def prox_state(self):
        return self.sim.prox_state()

    def prox_iters(self):
        return self.sim.prox_iters()

    def close(self):
        self.sim.close()


def run_scenario(adapter, scenario, dump=True, perturb=None):
    """Runs `frames` steps; returns dict of stacked per-frame arrays.  perturb = relative size of a random
    +-perturbation of the initial positions (used to measure the reference's own sensitivity)."""
    sc = scenario["scene"]
    x0 = np.asarray(sc.get("x_after_init", sc["x"]), dtype=np.float64)
    if perturb:
        sgn = np.random.default_rng(99).choice([-1.0, 1.0], size=x0.shape)
        x0 = x0 * (1.0 + perturb * sgn)
    if "x_after_init" in sc or perturb:
        adapter.set_x(x0)
    out = dict(x_it=[], z_it=[], u_it=[], x=[], v=[], prox_it=[])
    ev = scenario.get("events")
    for f in range(scenario["frames"]):
        if ev is not None:
            ev(f, adapter)
        if dump:
            xi, zi, ui, x, v = adapter.step_dump()
            out["x_it"].append(xi)
            out["z_it"].append(zi)
            out["u_it"].append(ui)
            pi = getattr(adapter, "last_prox_it", None)
            if pi is not None and pi.size:
                out["prox_it"].append(pi)
        else:
            x, v = adapter.step()
        out["x"].append(x)
        out["v"].append(v)
    res = {k: np.array(val) for k, val in out.items() if len(val)}
    ps = adapter.prox_state()
    if ps.size:
        res["prox_state"] = ps
        res["prox_iters"] = adapter.prox_iters()
    return res


# ---- scenes with events -------------------------------------------------------------------------------------
def _collide_scene():
    """plinkopony-shaped: ARAP tets + gravity + one CollisionForce over cylinders, a sphere and a floor."""
    sc = scenes.cube_scene(2, kind=TET_ARAP, stiffness=1e4, mass=10.0, dt=0.04, iters=13, stretch=None, name="collide")
    sc["x"] = sc["x"] + np.array([0.1, 0.45, 0.0])
    kinds = np.array([SHAPE_CYLINDER, SHAPE_CYLINDER, SHAPE_SPHERE, SHAPE_FLOOR], dtype=np.int32)
    params = np.array([[-0.3, -0.35, 0.0, 0.4], [0.65, -0.2, 0.0, 0.4], [0.1, -0.2, 0.1, 0.3], [0.0, -0.9, 0.0, 0.0]])
    sc["batches"].append(dict(type="collision", kinds=kinds, params=params, weight=32.0))
    return sc


def _anchor_scene():
    """poordillo-shaped: NH tets, uniform mass, gravity, MovingAnchors on two faces of a cube; one group is
    dragged with smooth_move (AnchorForce.hpp:33-40) and later released with weight 0 + recompute_weights()
    (poordillo.cpp:196-204)."""
    sc = scenes.cube_scene(2, kind=TET_NH, mu=1e5, lam=1e5, maxit=5, mass=1.0, dt=0.06, iters=10, stretch=None, name="anchors")
    x = sc["x"]
    sc["m"] = np.full(x.shape[0], 1.0 / x.shape[0])
    hand = np.where(x[:, 0] > 0.49)[0].astype(np.int32)
    foot = np.where(x[:, 0] < -0.49)[0].astype(np.int32)
    sc["batches"].append(dict(type="moving_anchors", idx=hand, pos=x[hand].copy(), weight=-1.0))
    sc["batches"].append(dict(type="moving_anchors", idx=foot, pos=x[foot].copy(), weight=-1.0))
    return sc


def _anchor_events(sc):
    hand_b, foot_b = 1, 2
    start = sc["batches"][hand_b]["pos"].copy()
    end = start + np.array([0.6, 0.1, 0.0])
    dt = sc["dt"]

    def smooth_move(t, t0, t1, a, b):
        if t < t0:
            return a
        r = (t - t0) / (t1 - t0)
        if r > 1.0:
            return b
        return a + (3.0 * r * r - 2.0 * r * r * r) * (b - a)

    def ev(frame, sim):
        t = frame * dt
        if frame < 5:
            sim.set_control_points(hand_b, pos=smooth_move(t, 0.0, 0.2, start, end))
        if frame == 5:  # release the hand: active = false, weight = 0, refactor
            n = len(sc["batches"][hand_b]["idx"])
            sim.set_control_points(hand_b, active=np.zeros(n, dtype=np.int32))
            sim.set_anchor_weights(hand_b, 0.0)
            sim.recompute_weights()
    return ev


def build_scenarios():
    S = {}
    S["singletet"] = dict(scene=scenes.singletet_scene(), frames=2)
    S["singlenode"] = dict(scene=scenes.singlenode_scene(), frames=4)
    S["cube3_nh"] = dict(scene=scenes.cube_scene(3, kind=TET_NH, seed=1), frames=3)
    S["cube2_nh10"] = dict(scene=scenes.cube_scene(2, kind=TET_NH, mu=3e4, lam=8e4, maxit=12, seed=5, name="cube2_nh10"), frames=3)
    S["cube2_stvk"] = dict(scene=scenes.cube_scene(2, kind=TET_STVK, mu=100.0, lam=100.0, mass=1.0, seed=2, name="cube2_stvk"), frames=3)
    S["cube2_arap"] = dict(scene=scenes.cube_scene(2, kind=TET_ARAP, stiffness=1e5, mass=10.0, seed=3, name="cube2_arap"), frames=3)
    vol = scenes.cube_scene(2, kind=TET_VOLUME, stiffness=1e4, mass=10.0, seed=4, name="cube2_vol")
    vol["batches"].insert(0, dict(type="tets", kind=TET_ARAP, idx=vol["batches"][0]["idx"], p0=1e3))
    S["cube2_vol"] = dict(scene=vol, frames=3)
    S["cloth6x4"] = dict(scene=scenes.cloth_scene(6, 4, springs=True, wind=(10.0, 0.0, 2.0), iters=10, name="cloth6x4"), frames=3)
    area = scenes.cloth_scene(5, 3, springs=False, wind=None, iters=8, name="cloth_area")
    area["batches"][0] = dict(type="tris", kind=TRI_AREA, idx=area["batches"][0]["idx"], stiffness=50.0, lmin=0.9, lmax=1.1, flag=3)
    S["cloth_area"] = dict(scene=area, frames=3)
    S["collide"] = dict(scene=_collide_scene(), frames=8)
    a = _anchor_scene()
    S["anchors"] = dict(scene=a, frames=8, events=_anchor_events(a))
    return S
