"""Explicit forces, CPU side: the product's wind arithmetic (csrc/elastic_math.h wind_triangle) and
its wavefront schedule (csrc/rest_state.cpp wind_wavefronts), compiled for the host by tests/hostcheck, against
 (a) the UNMODIFIED reference's goldens: x_bar = x + dt * v entering iteration 0 of every frame
     (System.cpp:37-46; goldens made with one OpenMP thread, the reference's deterministic meaning), bit for bit;
 (b) the serial loop: walking the wavefronts in any order inside a level gives the serial result, bit for bit.
"""
import ctypes as C
import os

import numpy as np
import pytest

import scenes
from admm_b200 import wind_project

HERE = os.path.dirname(os.path.abspath(__file__))
HC = os.path.join(HERE, "hostcheck", "libpipelinecheck.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _lib():
    if not os.path.exists(HC):
        pytest.skip("tests/hostcheck not built (run __graft_entry__.build())")
    L = C.CDLL(HC)
    L.hc_wind.argtypes = [C.c_int, C.c_int, _ip, _dp, C.c_double, _dp, _dp, C.c_int]
    return L


def _explicit(L, scene, x, v, serial=0):
    """System.cpp:37-39 with the product's wind."""
    dt = float(scene["dt"])
    depth = None
    for e in scene["explicit"]:
        if e["type"] == "gravity":
            v.reshape(-1, 3)[:] += dt * np.asarray(e["dir"], dtype=np.float64)
        else:
            tris = np.ascontiguousarray(e["tris"], dtype=np.int32)
            depth = L.hc_wind(x.size // 3, tris.size // 3, tris, np.ascontiguousarray(e["dir"], dtype=np.float64), dt, x, v, serial)
    return depth


@pytest.mark.parametrize("name", ["cloth6x4", "shipped_windyflag"])
def test_xbar_against_reference_goldens(name):
    L = _lib()
    scene = scenes.load_scene(os.path.join(HERE, "golden", f"{name}.scene.npz"))
    gold = np.load(os.path.join(HERE, "golden", f"{name}.ref.npz"))
    assert any(e["type"] == "wind" for e in scene["explicit"])
    dt = float(scene["dt"])
    x = np.ascontiguousarray(scene.get("x_after_init", scene["x"]), dtype=np.float64).reshape(-1).copy()
    v = np.zeros_like(x)
    frames = gold["x_it"].shape[0]
    depth = 0
    for f in range(frames):
        depth = _explicit(L, scene, x, v)
        xbar = x + dt * v
        assert np.array_equal(xbar, gold["x_it"][f, 0]), f"{name}: x_bar of frame {f} differs from the reference ({np.abs(xbar - gold['x_it'][f, 0]).max():.2e})"
        if f + 1 < frames:
            x, v = gold["x"][f].copy(), gold["v"][f].copy()
    print(f"{name}: x_bar bit-exact for {frames} frame(s); {depth} wavefronts")


def test_wavefronts_equal_serial_loop():
    L = _lib()
    rng = np.random.default_rng(11)
    sc = scenes.cloth_scene(12, 9, springs=False, wind=(3.0, 1.0, 4.0), iters=5, name="w")
    w = [e for e in sc["explicit"] if e["type"] == "wind"][0]
    tris = np.asarray(w["tris"], dtype=np.int32).reshape(-1, 3).copy()
    rng.shuffle(tris)
    tris = np.ascontiguousarray(np.concatenate([tris, tris[:5]]))
    n = sc["x"].shape[0]
    x = (np.asarray(sc["x"], dtype=np.float64) + 0.05 * rng.standard_normal((n, 3))).reshape(-1)
    v0 = rng.standard_normal(3 * n)
    d = np.array([3.0, 1.0, 4.0])
    va, vb, vc = v0.copy(), v0.copy(), v0.copy()
    depth = L.hc_wind(n, len(tris), tris, d, 0.04, x, va, 0)
    L.hc_wind(n, len(tris), tris, d, 0.04, x, vb, 1)
    wind_project(x, vc, tris, d, 0.04)
    assert 1 < depth < len(tris)
    assert np.array_equal(va, vb) and np.array_equal(va, vc)
    # and the order of the list matters (this is what the schedule must preserve)
    vd = v0.copy()
    L.hc_wind(n, len(tris), np.ascontiguousarray(tris[::-1]), d, 0.04, x, vd, 1)
    assert not np.array_equal(va, vd)
