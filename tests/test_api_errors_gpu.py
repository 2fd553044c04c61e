"""Error behaviour of the C ABI on the device (the reference returns false / throws; the ABI returns negative codes
with a message and never crosses the boundary with an exception)."""
import ctypes as C

import numpy as np
import pytest

import admm_b200
import scenes

pytestmark = pytest.mark.gpu


def _ctx():
    L = admm_b200.lib()
    h = C.c_void_p()
    assert L.admmb_create(0, C.byref(h)) == 0
    return L, h


def test_bad_node_index_is_rejected():
    L, h = _ctx()
    x = np.zeros(12)
    assert L.admmb_set_nodes(h, 4, x, np.ones(12)) == 0
    bad = np.array([[0, 1, 2, 7]], dtype=np.int32)
    assert L.admmb_add_tets(h, 0, 1, bad, 1.0, 0.0, 0.0, 0) == -1          # ADMMB_E_ARG
    assert b"out of range" in L.admmb_last_error(h)
    L.admmb_destroy(h)


def test_unequal_masses_per_coordinate_are_rejected():
    L, h = _ctx()
    m = np.ones(12)
    m[4] = 2.0
    assert L.admmb_set_nodes(h, 4, np.zeros(12), m) == -4                  # ADMMB_E_NUMERIC (A = A_n (x) I_3 needs it)
    L.admmb_destroy(h)


def test_calls_in_the_wrong_state():
    L, h = _ctx()
    x = np.zeros(3)
    assert L.admmb_step(h, 1, x, x.copy()) == -2                           # not finalized
    assert L.admmb_set_nodes(h, 1, x, np.ones(3)) == 0
    assert L.admmb_finalize(h, 0.04) == 0
    assert L.admmb_add_static_anchors(h, 1, np.zeros(1, dtype=np.int32), -1.0) == -2   # after finalize
    L.admmb_destroy(h)


def test_zero_mass_without_constraints_is_not_positive_definite():
    sc = scenes.cube_scene(2, kind=scenes.TET_ARAP, stretch=None)
    sc["m"] = np.zeros_like(sc["m"])
    # (the reference never checks SimplicialLDLT's info, System.cpp:140; the direct solver here reports the failure.
    #  PCG only sees a positive diagonal and cannot tell at setup time.)
    with pytest.raises(admm_b200.AdmmError, match="positive definite|non-positive"):
        admm_b200.System(sc)


def test_nonpositive_timestep_falls_back_to_default():
    """System.cpp:103-107: a timestep <= 0 is replaced by 0.04 s."""
    sc = scenes.cube_scene(2, kind=scenes.TET_ARAP, stretch=None)
    a = admm_b200.System(dict(sc, dt=-1.0))
    b = admm_b200.System(dict(sc, dt=0.04))
    a.dt = b.dt = 0.04
    a.step()
    b.step()
    assert np.allclose(a.m_x, b.m_x, rtol=0, atol=1e-13)


def test_empty_batches_and_force_free_system():
    sc = scenes.singlenode_scene()
    sc["batches"] = [dict(type="tets", kind=0, idx=np.zeros((0, 4), dtype=np.int32), p0=1.0)]
    s = admm_b200.System(sc)
    s.step()
    assert abs(s.m_x[1] - (-9.800000190734863)) < 1e-12
    s.close()


def test_state_roundtrip_checkpoint_resume():
    """x, v, u and the L-BFGS state define a resume point: restoring them reproduces the next frame."""
    sc = scenes.cube_scene(3, kind=scenes.TET_NH, seed=5)
    a = admm_b200.System(sc)
    a.set_x(sc["x_after_init"])
    for _ in range(2):
        a.step()
    x, v, u, prox = a.m_x.copy(), a.m_v.copy(), a.u, a.prox_state()
    a.step()
    b = admm_b200.System(sc)
    b.m_x[:], b.m_v[:] = x, v
    b.set_state(admm_b200.STATE_U, u)
    b.set_state(admm_b200.STATE_PROX, prox)
    b.step()
    err = np.linalg.norm(a.m_x - b.m_x) / np.linalg.norm(a.m_x)
    assert err < 1e-3     # not bit-exact: the solve's atomics are unordered and the NH prox is chaotic at this level
    a.close()
    b.close()


def test_registered_host_buffers_give_the_same_results():
    """admmb_register_host_buffer only changes how x / v travel (direct DMA instead of staging), not what is computed;
    buffers that were never registered -- or other arrays passed later -- keep working."""
    sc = scenes.cube_scene(3, kind=scenes.TET_ARAP, iters=6)
    a = admm_b200.System(sc, pin_host=True)
    b = admm_b200.System(sc, pin_host=False)
    for s in (a, b):
        s.set_x(sc["x_after_init"])
    for _ in range(3):
        a.step()
        b.step()
    assert np.abs(a.m_x - b.m_x).max() <= 1e-12 and np.abs(a.m_v - b.m_v).max() <= 1e-11
    # an unregistered pair of arrays on the pinned system: staged path
    x2, v2 = a.m_x.copy(), a.m_v.copy()
    assert a.L.admmb_step(a.h, 6, x2, v2) == 0
    b.step()
    assert np.abs(x2 - b.m_x).max() <= 1e-12
    # registering twice is a no-op, unregistering an unknown pointer is an argument error
    p = a.m_x.ctypes.data_as(C.c_void_p)
    assert a.L.admmb_register_host_buffer(a.h, p, a.m_x.nbytes) == 0
    assert a.L.admmb_unregister_host_buffer(a.h, x2.ctypes.data_as(C.c_void_p)) == -1
    assert a.L.admmb_unregister_host_buffer(a.h, p) == 0
    a.step()   # m_x now staged again, m_v still direct
    a.close()
    b.close()


def test_async_stepping_of_several_contexts_matches_sequential_stepping():
    """Scene ensembles keep several contexts in flight on one GPU (admmb_step_resident_async + admmb_sync): the
    result of every scene must be what it is when stepped alone."""
    scs = [scenes.cube_scene(3, kind=scenes.TET_ARAP, iters=6, seed=10 + i) for i in range(4)]
    alone = []
    for sc in scs:
        s = admm_b200.System(sc)
        s.set_x(sc["x_after_init"])
        s.upload()
        s.step_resident(frames=5)
        s.download()
        alone.append(s.m_x.copy())
        s.close()
    sims = [admm_b200.System(sc) for sc in scs]
    for s, sc in zip(sims, scs):
        s.set_x(sc["x_after_init"])
        s.upload()
    for s in sims:
        s.step_resident_async(frames=5)
    for s in sims:
        s.sync()
        assert s.last_region_ms() > 0.0
    for s, ref in zip(sims, alone):
        s.download()
        assert np.abs(s.m_x - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        s.close()


def test_deterministic_solve_is_bit_reproducible_and_agrees_with_the_default():
    """admmb_set_deterministic: two runs on the same input are bit-identical (the default accumulates the solve with
    floating-point atomics, reproducible only to rounding), and both modes agree to rounding where the dynamics do not
    amplify it (ARAP; the NeoHookean line search does)."""

    def run(sc, det):
        s = admm_b200.System(sc, deterministic=det)
        s.set_x(sc["x_after_init"])
        xs = []
        for _ in range(4):
            s.step()
            xs.append(s.m_x.copy())
        s.close()
        return np.array(xs)

    nh = scenes.cube_scene(8, kind=scenes.TET_NH, seed=3)
    a, b = run(nh, True), run(nh, True)
    assert np.array_equal(a, b), "deterministic mode is not bit-reproducible"
    arap = scenes.cube_scene(8, kind=scenes.TET_ARAP, seed=3)
    c, d = run(arap, True), run(arap, False)
    assert np.abs(c - d).max() <= 1e-12 * np.abs(c).max()


def test_check_finite_reports_divergence():
    """admmb_set_check_finite: once a frame produces a non-finite position the step fails with ADMMB_E_NUMERIC (the
    reference's step() always returns true, System.cpp:74, so this is off by default)."""
    sc = scenes.cloth_scene(6, 5, springs=False, wind=(3.0, 1.0, 4.0), iters=6, name="diverge")
    a = admm_b200.System(sc, iters=0)      # explicit wind drag without the implicit solve is unstable: NaN within ~10 frames
    assert a.L.admmb_set_check_finite(a.h, 1) == 0
    failed_at = None
    for f in range(40):
        rc = a.L.admmb_step(a.h, 0, a.m_x, a.m_v)
        if rc != 0:
            failed_at = f
            assert rc == -4 and b"not finite" in a.L.admmb_last_error(a.h)
            break
    assert failed_at is not None and not np.isfinite(a.m_x).all()
    a.close()
    b = admm_b200.System(sc, iters=0)      # default: no check, the call keeps returning 0 like the reference
    for f in range(failed_at + 1):
        assert b.L.admmb_step(b.h, 0, b.m_x, b.m_v) == 0
    b.close()


def test_step_async_with_host_buffers_matches_step():
    """admmb_step_async + admmb_sync = admmb_step, for page-locked and for plain host buffers, several scenes in flight."""
    scs = [scenes.cube_scene(3, kind=scenes.TET_ARAP, iters=6, seed=20 + i) for i in range(3)]
    ref = []
    for sc in scs:
        s = admm_b200.System(sc)
        s.set_x(sc["x_after_init"])
        for _ in range(3):
            s.step()
        ref.append((s.m_x.copy(), s.m_v.copy()))
        s.close()
    for pin in (True, False):
        sims = [admm_b200.System(sc, pin_host=pin) for sc in scs]
        for s, sc in zip(sims, scs):
            s.set_x(sc["x_after_init"])
        for _ in range(3):
            for s in sims:
                s.step_async()
            for s in sims:
                s.sync()
        for s, (x, v) in zip(sims, ref):
            assert np.abs(s.m_x - x).max() <= 1e-12 and np.abs(s.m_v - v).max() <= 1e-10
        # a second step_async before sync is a state error
        assert sims[0].L.admmb_step_async(sims[0].h, 6, sims[0].m_x, sims[0].m_v) == 0
        assert sims[0].L.admmb_step_async(sims[0].h, 6, sims[0].m_x, sims[0].m_v) == -2
        sims[0].sync()
        for s in sims:
            s.close()


def test_pending_async_step_blocks_calls_that_would_overwrite_its_staging():
    """admmb_step_async parks the step's results in the context's staging areas until admmb_sync: every call that uses the
    same areas must refuse to run in between (ADMMB_E_STATE) instead of silently corrupting the pending x / v."""
    sc = scenes.cube_scene(3, kind=scenes.TET_ARAP, seed=3)
    sim = admm_b200.System(sc)
    sim.set_x(sc["x_after_init"])
    sim.step()
    ref = admm_b200.System(sc)
    ref.set_x(sc["x_after_init"])
    ref.step()
    ref.step()
    L, h = sim.L, sim.h
    sim.step_async()
    f32 = np.zeros(sim.n3, dtype=np.float32)
    tmp = np.zeros(sim.n3)
    assert L.admmb_download_x_f32(h, f32) == -2
    assert L.admmb_debug_local_step(h, tmp) == -2
    assert L.admmb_get_state(h, admm_b200.STATE_X, tmp) == -2
    assert L.admmb_step(h, 1, tmp, tmp.copy()) == -2
    assert L.admmb_recompute_weights(h) == -2
    assert b"admmb_sync" in L.admmb_last_error(h)
    sim.sync()
    assert np.linalg.norm(sim.m_x - ref.m_x) <= 1e-12 * np.linalg.norm(ref.m_x)
    assert L.admmb_download_x_f32(h, f32) == 0
    sim.close()
    ref.close()


def test_failed_refactorisation_makes_the_context_refuse_to_step_until_repaired():
    sc = scenes.cube_scene(2, kind=scenes.TET_ARAP, seed=3)
    sim = admm_b200.System(sc)
    T = sc["batches"][0]["idx"].shape[0]
    w0 = sim.get_batch_weights(sim.batch_ids[0], T)
    sim.set_batch_weights(sim.batch_ids[0], np.full(T, np.nan))
    with pytest.raises(admm_b200.AdmmError):
        sim.recompute_weights()                   # NaN weights: the factorisation reports a pivot failure
    x = sim.m_x.copy()
    assert sim.L.admmb_step(sim.h, 1, x, x.copy()) == -2
    assert b"unusable" in sim.L.admmb_last_error(sim.h)
    sim.set_batch_weights(sim.batch_ids[0], w0)
    sim.recompute_weights()                       # repaired
    sim.step()
    assert np.isfinite(sim.m_x).all()
    sim.close()


def test_disabled_explicit_force_is_skipped():
    sc = scenes.cube_scene(2, kind=scenes.TET_ARAP, stretch=None)
    a = admm_b200.System(sc, iters=0)
    b = admm_b200.System(sc, iters=0)
    assert len(a.gravity_ids) == 1
    assert a.L.admmb_enable_explicit(a.h, a.gravity_ids[0], 0) == 0
    assert a.L.admmb_enable_explicit(a.h, 99, 0) == -1
    a.step()
    b.step()
    assert np.array_equal(a.m_v, np.zeros_like(a.m_v))       # no gravity: nothing moves
    assert np.abs(b.m_v).max() > 0
    a.L.admmb_enable_explicit(a.h, a.gravity_ids[0], 1)
    a.step()
    assert np.array_equal(a.m_v, b.m_v)                      # first frame of gravity from rest, as b's
    a.close()
    b.close()
