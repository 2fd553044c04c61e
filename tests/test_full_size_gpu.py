"""Parity at BASELINE.json's full size (cube N=55: 998 250 NeoHookean tets, the benchmark workload).

The reference cannot step this mesh in test time (its initialize() alone takes ~44 min), but the local step of a tet
depends only on the positions of its four nodes, its u and its optimiser state.  So a random sample of tets of the
full-size device run is replayed through the UNMODIFIED reference as a "tet soup" (every sampled tet a separate
4-node component with the same rest shape): positions entering every ADMM iteration are taken from the device's dumps,
the reference steps one iteration, and z, u must agree BIT FOR BIT -- from rest and in the benchmark's steady regime.
The global step at full size is checked by an independent algorithm: the Jacobi-PCG solve of the same system."""
import numpy as np
import pytest

import admm_b200
import scenes
from util import have_ref, rel_l2

pytestmark = pytest.mark.gpu

N = 55
SAMPLE = 4096


def _soup(sc, tets, sample):
    x = np.asarray(sc["x"], dtype=np.float64)[tets[sample]].reshape(-1, 3)
    b = dict(sc["batches"][0])
    b["idx"] = np.arange(4 * len(sample), dtype=np.int32).reshape(-1, 4)
    return dict(name="soup", dt=sc["dt"], iters=1, x=x, m=np.ones(x.shape[0]), batches=[b], explicit=[], extra={})


def test_sampled_tets_of_the_full_size_run_match_the_reference_bit_for_bit():
    if not have_ref():
        pytest.skip("oracle/_ref/libadmm_ref.so not present on this box")
    from oracle import ref
    sc = scenes.cube_scene(N)
    tets = np.asarray(sc["batches"][0]["idx"]).reshape(-1, 4)
    T = tets.shape[0]
    assert T == 998250
    sample = np.sort(np.random.default_rng(7).choice(T, SAMPLE, replace=False))
    rows = (9 * sample[:, None] + np.arange(9)[None, :]).reshape(-1)
    rs = ref.RefSystem(_soup(sc, tets, sample), iters=1)
    sim = admm_b200.System(sc)
    sim.set_x(sc["x_after_init"])

    def replay(xi, zi, ui, label):
        exact = 0
        for k in range(xi.shape[0]):
            rs.set_x(xi[k].reshape(-1, 3)[tets[sample]].reshape(-1))
            rs.set_v(np.zeros(12 * SAMPLE))           # x_bar = x + dt * 0 = x exactly: the soup sees the device's curr_x
            rs.step()
            ok = np.array_equal(rs.z, zi[k][rows]) and np.array_equal(rs.u, ui[k][rows])
            exact += int(ok)
            assert ok, (f"{label}, iteration {k}: sampled z / u differ from the reference "
                        f"(z {rel_l2(zi[k][rows], rs.z):.1e}, u {rel_l2(ui[k][rows], rs.u):.1e})")
        return exact

    # from rest: frame 0, iterations 0..2 (u = 0, optimiser state = initial on both sides)
    xi, zi, ui = sim.step_dump(iters=3)
    n0 = replay(xi, zi, ui, "frame 0")
    # steady state: 19 more frames on the device, then hand u and the optimiser state of the sampled tets to the reference
    sim.step_resident(frames=19)
    u_all, prox_all = sim.u, sim.prox_state()
    sim.download()
    rs.set_u(u_all[rows])
    rs.set_prox_state(prox_all[sample])
    xi, zi, ui = sim.step_dump(iters=2)
    n1 = replay(xi, zi, ui, "frame 20")
    print(f"full size ({T} tets): {SAMPLE} sampled tets bit-exact against the reference in {n0} + {n1} iterations "
          f"(frame 0 from rest; frame 20, steady state)")
    rs.close()
    sim.close()


def test_full_size_direct_solve_agrees_with_pcg():
    """Same right-hand side, two independent algorithms (supernodal factor + tile solve vs Jacobi-PCG on A_n)."""
    sc = scenes.cube_scene(N)
    a = admm_b200.System(sc, solver=admm_b200.SOLVER_DIRECT)
    b = admm_b200.System(sc, solver=admm_b200.SOLVER_PCG, cg_tol=1e-13, cg_max_iters=20000)
    x0 = np.asarray(sc["x_after_init"], dtype=np.float64).reshape(-1)
    for s in (a, b):
        s.debug_local_step(x0)          # identical z, u on both (bit-exact local step)
        s.debug_global_step(x0)
    err = rel_l2(a.x_iter, b.x_iter)
    print(f"full size: direct vs PCG solution rel-L2 {err:.1e}")
    assert np.isfinite(a.x_iter).all()
    assert err <= 1e-9
    a.close()
    b.close()
