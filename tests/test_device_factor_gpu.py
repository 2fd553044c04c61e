"""Device backend of the factorisation (csrc/front_gpu.cu): the large fronts of the multifrontal
Cholesky are factored on the B200 (cuSOLVER / cuBLAS, FP64).  Forced down to small fronts here (ADMMB_GPU_FRONT_MIN) so
that the unit-sized scenes exercise it; the results must agree with the all-host factorisation to rounding and with the
reference's goldens."""
import os

import numpy as np
import pytest

import admm_b200
import scenes

pytestmark = pytest.mark.gpu


def _run(sc, frames, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        sim = admm_b200.System(sc)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    info = sim.info()
    sim.set_x(sc["x_after_init"])
    xs = []
    for _ in range(frames):
        sim.step()
        xs.append(sim.m_x.copy())
    sim.close()
    return np.array(xs), info


@pytest.mark.parametrize("N", [6, 12])
def test_device_fronts_match_host_factorisation(N):
    sc = scenes.cube_scene(N, kind=scenes.TET_ARAP, iters=8)
    xh, ih = _run(sc, 3, {"ADMMB_HOST_FACTOR": "1"})
    xd, idv = _run(sc, 3, {"ADMMB_HOST_FACTOR": "0", "ADMMB_GPU_FRONT_MIN": "48"})
    assert ih["device_fronts"] == 0
    assert idv["device_fronts"] > 0, "the device backend did not run (cuSOLVER / cuBLAS not loadable?)"
    assert idv["nnz_L"] == ih["nnz_L"] and idv["n_levels"] == ih["n_levels"]
    err = np.abs(xd - xh).max() / np.abs(xh).max()
    print(f"N={N}: {idv['device_fronts']} fronts on the device of {idv['n_supernodes']} supernodes; x vs host factor {err:.1e}")
    assert err <= 1e-12


def test_indefinite_front_is_reported_from_the_device(monkeypatch):
    """A massless, unconstrained node makes A singular: the device potrf must report it like the host path does."""
    import ctypes as C
    monkeypatch.setenv("ADMMB_GPU_FRONT_MIN", "16")
    L = admm_b200.lib()
    h = C.c_void_p()
    assert L.admmb_create(0, C.byref(h)) == 0
    x, tets = scenes.kuhn_cube(3)
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    m = np.ones(3 * n)
    m[:] = -1.0     # negative masses: M + dt^2 D^T W^2 D is indefinite
    assert L.admmb_set_nodes(h, n, x.reshape(-1).copy(), m) == 0
    t = np.ascontiguousarray(tets, dtype=np.int32)
    assert L.admmb_add_tets(h, 0, t.shape[0], t, 1e-3, 0.0, 0.0, 0) >= 0
    rc = L.admmb_finalize(h, 0.04)
    assert rc == -4 and b"positive definite" in L.admmb_last_error(h)
    L.admmb_destroy(h)
