"""Explicit forces on the device: ExplicitForce over all nodes or a subset and WindForce
(A/src/system/ExplicitForce.cpp:29-98) applied inside admmb_step, against the same forces applied to m_v by the
caller with the host restatement (admm_b200.wind_project = the reference on one OpenMP thread, itself pinned by the
windyflag goldens; tests/test_explicit_cpu.py pins the product's arithmetic against the reference's goldens directly).

A step with 0 ADMM iterations returns x = x_bar = x + dt * v_explicit and v = v_explicit (System.cpp:37-48,70-72), which
isolates the explicit stage: there the two must agree BIT FOR BIT (same triangle order per node, one rounding per
operation).  With iterations the global solve accumulates with floating-point atomics (order varies run to run at the
1e-16 level), so full steps are compared to 1e-12.
"""
import os

import numpy as np
import pytest

import admm_b200
import scenes
from scenarios import build_shipped

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _run(scene, frames, host_explicit, iters=0):
    sim = admm_b200.System(scene, host_explicit=host_explicit, iters=iters)
    xs, vs = [], []
    for _ in range(frames):
        if host_explicit:
            sim.apply_host_explicit()
        sim.step()
        xs.append(sim.m_x.copy())
        vs.append(sim.m_v.copy())
    sim.close()
    return np.array(xs), np.array(vs)


def test_device_wind_equals_host_wind_synthetic_cloth():
    sc = scenes.cloth_scene(9, 7, springs=True, wind=(10.0, 0.0, 2.0), iters=10, name="cloth_wind")
    # (without the implicit solve the explicit drag is unstable: keep the isolated run to 3 frames)
    xd, vd = _run(sc, 3, False)
    xh, vh = _run(sc, 3, True)
    assert np.isfinite(xd).all() and np.array_equal(xd, xh) and np.array_equal(vd, vh)
    assert np.abs(vd[-1]).max() > 1e-3  # the wind did something
    xd, vd = _run(sc, 6, False, iters=10)
    xh, vh = _run(sc, 6, True, iters=10)
    assert np.abs(xd - xh).max() <= 1e-12 and np.abs(vd - vh).max() <= 1e-11


def test_device_wind_equals_host_wind_shipped_flag():
    S = build_shipped(GOLD)
    if "windyflag" not in S:
        pytest.skip("no windyflag fixture")
    sc = S["windyflag"]["scene"]
    xd, vd = _run(sc, 3, False)
    xh, vh = _run(sc, 3, True)
    assert np.isfinite(xd).all() and np.array_equal(xd, xh) and np.array_equal(vd, vh)
    xd, vd = _run(sc, 3, False, iters=30)
    xh, vh = _run(sc, 3, True, iters=30)
    assert np.abs(xd - xh).max() <= 1e-12


def test_wind_with_shuffled_and_repeated_triangles():
    """Triangle order defines the result (serial dependency through v): a shuffled list with triangles listed twice
    must follow the reference's loop literally."""
    sc = scenes.cloth_scene(6, 5, springs=False, wind=(3.0, 1.0, 4.0), iters=6, name="cloth_wind2")
    w = [e for e in sc["explicit"] if e["type"] == "wind"][0]
    tris = np.asarray(w["tris"]).reshape(-1, 3).copy()
    rng = np.random.default_rng(5)
    rng.shuffle(tris)
    tris = np.concatenate([tris, tris[:3], tris[-2:]]).astype(np.int32)
    w["tris"] = tris
    xd, vd = _run(sc, 3, False)
    xh, vh = _run(sc, 3, True)
    assert np.array_equal(xd, xh) and np.array_equal(vd, vh)


def test_explicit_subset_with_duplicates_and_order():
    """ExplicitForce with `indices` (a node listed twice gets the increment twice), mixed with wind and gravity."""
    sc = scenes.cloth_scene(6, 5, springs=False, wind=(3.0, 1.0, 4.0), iters=6, name="cloth_subset")
    n = sc["x"].size // 3
    sub = dict(type="gravity", dir=np.array([0.3, 0.7, -0.2]), indices=np.array([1, 5, 5, n - 1, 2, 5], dtype=np.int32))
    sc["explicit"] = [sub] + list(sc["explicit"]) + [dict(type="gravity", dir=np.array([0.0, -1.0, 0.5]))]
    xd, vd = _run(sc, 3, False)
    xh, vh = _run(sc, 3, True)
    assert np.array_equal(xd, xh) and np.array_equal(vd, vh)


def test_wind_direction_can_change_between_steps():
    sc = scenes.cloth_scene(6, 5, springs=False, wind=(3.0, 1.0, 4.0), iters=6, name="cloth_dir")
    a = admm_b200.System(sc)
    b = admm_b200.System(sc, host_explicit=True)
    wid = [i for i, e in enumerate(sc["explicit"]) if e["type"] == "wind"][0]
    for f in range(4):
        d = np.array([3.0 + f, 1.0, 4.0 - f])
        a.set_gravity(a.gravity_ids[wid], d)
        b.host_explicit[wid] = dict(b.host_explicit[wid], dir=d)
        b.apply_host_explicit()
        a.step(0)
        b.step(0)
        assert np.array_equal(a.m_x, b.m_x) and np.array_equal(a.m_v, b.m_v)
    a.close()
    b.close()


def test_explicit_argument_errors():
    import ctypes as C
    L = admm_b200.lib()
    h = C.c_void_p()
    assert L.admmb_create(0, C.byref(h)) == 0
    d = np.array([0.0, -9.8, 0.0])
    tri = np.array([0, 1, 2], dtype=np.int32)
    assert L.admmb_add_wind(h, 1, tri, d) == -2                                  # before set_nodes: ADMMB_E_STATE
    assert L.admmb_set_nodes(h, 3, np.array([0., 0, 0, 1, 0, 0, 0, 1, 0]), np.ones(9)) == 0
    assert L.admmb_add_wind(h, 1, np.array([0, 1, 3], dtype=np.int32), d) == -1  # index out of range
    assert L.admmb_add_explicit_subset(h, 0, tri, d) == -1                       # empty subset means "all" upstream
    assert L.admmb_add_wind(h, 1, tri, d) == 0
    assert L.admmb_add_explicit_subset(h, 2, tri, d) == 1
    assert L.admmb_add_static_anchors(h, 1, np.array([0], dtype=np.int32), -1.0) >= 0
    assert L.admmb_finalize(h, 0.04) == 0
    assert L.admmb_add_wind(h, 1, tri, d) == -2                                  # after finalize
    assert L.admmb_set_gravity(h, -1, d) == 2                                    # plain gravity may still be added
    assert L.admmb_set_gravity(h, 7, d) == -1
    x = np.array([0., 0, 0, 1, 0, 0, 0, 1, 0])
    v = np.zeros(9)
    assert L.admmb_step(h, 3, x, v) == 0
    assert np.isfinite(x).all() and np.abs(v).max() > 0
    L.admmb_destroy(h)
