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


def run_scenario(adapter, scenario, dump=True, perturb=None, seed=99):
    """Runs `frames` steps; returns dict of stacked per-frame arrays.  perturb = relative size of a random
    +-perturbation of the initial positions (used to measure the reference's own sensitivity)."""
    sc = scenario["scene"]
    x0 = np.asarray(sc.get("x_after_init", sc["x"]), dtype=np.float64)
    if perturb:
        # relative AND absolute (a planar cloth has z == 0 everywhere: a purely relative perturbation would leave the
        # out-of-plane direction, the one the wind excites, untouched)
        rng = np.random.default_rng(seed)
        sgn = rng.choice([-1.0, 1.0], size=x0.shape)
        x0 = x0 * (1.0 + perturb * sgn) + perturb * np.abs(x0).max() * rng.choice([-1.0, 1.0], size=x0.shape)
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
    fung = scenes.cloth_scene(5, 3, springs=False, wind=None, iters=8, name="cloth_fung")
    fung["batches"][0] = dict(type="tris", kind=scenes.TRI_FUNG, idx=fung["batches"][0]["idx"], stiffness=30.0, lmin=0.0, lmax=0.0, flag=0)
    fung["batches"] = [fung["batches"][0], fung["batches"][2]]   # FungTriangle + static anchors
    S["cloth_fung"] = dict(scene=fung, frames=3)
    S["collide"] = dict(scene=_collide_scene(), frames=8)
    a = _anchor_scene()
    S["anchors"] = dict(scene=a, frames=8, events=_anchor_events(a))
    return S


# ---- the reference's four shipped scenes (BASELINE.json configs[0..3]) ----------------------------------------
SHIPPED = ("bunnyexpand", "windyflag", "poordillo", "plinkopony")


def parse_exported_scene(path, name):
    """Text written by oracle/scene_export.cpp (the reference's own scene layer) -> scene dictionary.  Maximal runs
    of equal forces become batches, in force order."""
    tok = open(path).read().split()
    pos = [0]

    def nxt(cast=str):
        v = tok[pos[0]]
        pos[0] += 1
        return cast(v)

    sc = dict(name=name, batches=[], explicit=[])
    runs = []
    while pos[0] < len(tok):
        key = nxt()
        if key == "settings":
            sc["dt"], sc["iters"] = nxt(float), nxt(int)
        elif key == "nodes":
            n = nxt(int)
            a = np.array([nxt(float) for _ in range(4 * n)]).reshape(n, 4)
            sc["x"], sc["m"] = a[:, :3].copy(), a[:, 3].copy()
        elif key == "x_after":
            n = nxt(int)
            sc["x_after_init"] = np.array([nxt(float) for _ in range(3 * n)]).reshape(n, 3)
        elif key == "forces":
            nf = nxt(int)
            for _ in range(nf):
                t = nxt()
                if t == "tet":
                    kind = nxt(int)
                    idx = [nxt(int) for _ in range(4)]
                    p0, p1, p2, maxit = nxt(float), nxt(float), nxt(float), nxt(int)
                    runs.append((("tets", kind, p0, p1, p2, maxit), idx))
                elif t == "tri":
                    kind = nxt(int)
                    idx = [nxt(int) for _ in range(3)]
                    st, lo, hi, flag = nxt(float), nxt(float), nxt(float), nxt(int)
                    runs.append((("tris", kind, st, lo, hi, flag), idx))
                elif t == "bend":
                    idx = [nxt(int) for _ in range(4)]
                    runs.append((("bends", nxt(float)), idx))
                elif t == "spring":
                    idx = [nxt(int) for _ in range(2)]
                    runs.append((("springs", nxt(float)), idx))
                elif t == "sanchor":
                    i, w = nxt(int), nxt(float)
                    runs.append((("static_anchors", w), [i]))
                elif t == "manchor":
                    i, w = nxt(int), nxt(float)
                    p = [nxt(float) for _ in range(3)]
                    runs.append((("moving_anchors", w), [i] + p))
                elif t == "collision":
                    w, ns = nxt(float), nxt(int)
                    kinds, par = [], []
                    for _ in range(ns):
                        assert nxt() == "shape"
                        kinds.append(nxt(int))
                        par.append([nxt(float) for _ in range(4)])
                    runs.append((("collision", w), (kinds, par)))
                else:
                    raise ValueError(t)
        elif key == "explicit":
            ne = nxt(int)
            for _ in range(ne):
                t = nxt()
                d = np.array([nxt(float) for _ in range(3)])
                if t == "gravity":
                    sc["explicit"].append(dict(type="gravity", dir=d))
                else:
                    nt = nxt(int)
                    tr = np.array([nxt(int) for _ in range(3 * nt)], dtype=np.int32).reshape(nt, 3)
                    sc["explicit"].append(dict(type="wind", dir=d, tris=tr))
        else:
            raise ValueError(key)
    # group maximal runs
    i = 0
    while i < len(runs):
        key = runs[i][0]
        j = i
        while j < len(runs) and runs[j][0] == key and key[0] != "collision":
            j += 1
        j = max(j, i + 1)
        items = [r[1] for r in runs[i:j]]
        t = key[0]
        if t == "tets":
            sc["batches"].append(dict(type="tets", kind=key[1], idx=np.array(items, dtype=np.int32), p0=key[2], p1=key[3], p2=key[4], maxit=key[5]))
        elif t == "tris":
            sc["batches"].append(dict(type="tris", kind=key[1], idx=np.array(items, dtype=np.int32), stiffness=key[2], lmin=key[3], lmax=key[4], flag=key[5]))
        elif t == "bends":
            sc["batches"].append(dict(type="bends", idx=np.array(items, dtype=np.int32), stiffness=key[1]))
        elif t == "springs":
            sc["batches"].append(dict(type="springs", idx=np.array(items, dtype=np.int32), stiffness=key[1]))
        elif t == "static_anchors":
            sc["batches"].append(dict(type="static_anchors", idx=np.array([it[0] for it in items], dtype=np.int32), weight=key[1]))
        elif t == "moving_anchors":
            sc["batches"].append(dict(type="moving_anchors", idx=np.array([it[0] for it in items], dtype=np.int32),
                                      pos=np.array([it[1:] for it in items], dtype=np.float64), weight=key[1]))
        elif t == "collision":
            kinds, par = items[0]
            sc["batches"].append(dict(type="collision", kinds=np.array(kinds, dtype=np.int32), params=np.array(par, dtype=np.float64), weight=key[1]))
        i = j
    return sc


def _poordillo_events(sc):
    """Headless stand-in for the GUI interaction of samples/poordillo: the hand sphere's control points are dragged
    from (.6,.8,.5) towards (2.6,.8,.5) with helper::smooth_move (poordillo.cpp:51-58, AnchorForce.hpp:33-40); at
    frame 20 the hand is released as the H key does (poordillo.cpp:196-204: active = false, weight = 0,
    recompute_weights())."""
    mov = [i for i, b in enumerate(sc["batches"]) if b["type"] == "moving_anchors"]
    x = sc["x"]
    hand_c = np.array([.6, .8, .5], dtype=np.float32).astype(np.float64)
    # the exporter pushed hand anchors first, then foot anchors; they form ONE run (same class), so split by position
    b = sc["batches"][mov[0]]
    is_hand = np.linalg.norm(b["pos"] - hand_c, axis=1) < 0.2 + 1e-6
    start = b["pos"].copy()
    end = start + np.where(is_hand[:, None], np.array([2.0, 0.0, 0.0]), 0.0)
    dt = sc["dt"]

    def ev(frame, sim):
        t = frame * dt
        if frame < 20:
            r = min(max(t / 1.2, 0.0), 1.0)
            sim.set_control_points(mov[0], pos=start + (3.0 * r * r - 2.0 * r * r * r) * (end - start))
        if frame == 20:
            act = np.where(is_hand, 0, 1).astype(np.int32)
            sim.set_control_points(mov[0], active=act)
            sim.set_anchor_weights(mov[0], np.where(is_hand, 0.0, 1000.0))
            sim.recompute_weights()
    return ev


SHIPPED_FRAMES = dict(bunnyexpand=12, windyflag=8, poordillo=26, plinkopony=20)  # windyflag: the explicit wind amplifies rounding noise ~3x per frame (5.8e-10 after 12 frames)


# BASELINE.json configs 1 and 3 name material VARIANTS of two shipped scenes ("poordillo ... ARAP and StVK variants",
# "bunnyexpand ... NeoHookean"; SURVEY 8d: "run both"): the same meshes, anchors and events with the tet force swapped the
# way editing the `type` attribute of the scene XML does (ForceBuilder.cpp:276-446: LinearTetStrain stiffness, StVKTet /
# NeoHookeanTet mu, lambda, max_iterations).
VARIANTS = {
    "poordillo_arap": ("poordillo", dict(kind=TET_ARAP, p0=1e5, p1=0.0, p2=0.0, maxit=0)),
    "poordillo_stvk": ("poordillo", dict(kind=TET_STVK)),
    "bunnyexpand_nh": ("bunnyexpand", dict(kind=TET_NH)),
}


def shipped_variant(sc, change):
    """Copy of a shipped scene with every tet batch's material replaced (mu / lambda / max_iterations kept unless given)."""
    out = dict(sc)
    out["batches"] = []
    for b in sc["batches"]:
        b = dict(b)
        if b["type"] == "tets":
            b.update(change)
        out["batches"].append(b)
    return out


def build_shipped(golden_dir, variants=True):
    """The four shipped scenes as exported by the reference's scene layer (tests/golden/shipped_*.scene.npz), plus the
    material variants BASELINE.json names."""
    import os
    S = {}
    for name in SHIPPED:
        p = os.path.join(golden_dir, f"shipped_{name}.scene.npz")
        if not os.path.exists(p):
            continue
        sc = scenes.load_scene(p)
        S[name] = dict(scene=sc, frames=SHIPPED_FRAMES[name])
        if name == "poordillo":
            S[name]["events"] = _poordillo_events(sc)
    if variants:
        for vname, (base, change) in VARIANTS.items():
            if base not in S:
                continue
            sc = shipped_variant(S[base]["scene"], change)
            sc["name"] = vname
            S[vname] = dict(scene=sc, frames=SHIPPED_FRAMES[base])
            if base == "poordillo":
                S[vname]["events"] = _poordillo_events(sc)
    return S


# 100-frame trajectories (north_star: "trajectories over 100 frames within 1e-6 for collision-free scenes"): scenes whose
# reference run is reproducible over that horizon.  No user interaction; every 10th frame is kept in the golden file.
def build_long(golden_dir):
    S = build_shipped(golden_dir)
    L = {}
    if "poordillo_arap" in S:
        L["poordillo_arap_100"] = dict(scene=S["poordillo_arap"]["scene"], frames=100)
    if "plinkopony" in S:   # the pony without its pegs: collision-free ARAP under gravity
        sc = dict(S["plinkopony"]["scene"])
        sc["batches"] = [b for b in sc["batches"] if b["type"] != "collision"]
        sc["name"] = "plinkopony_nocontact"
        L["plinkopony_nocontact_100"] = dict(scene=sc, frames=100)
    if "windyflag" in S:
        L["windyflag_100"] = dict(scene=S["windyflag"]["scene"], frames=100)
    return L
