"""The reference's four shipped scenes (BASELINE.json configs[0..3]: bunnyexpand, windyflag, poordillo, plinkopony) and the
material variants BASELINE.json names (poordillo ARAP / StVK, bunnyexpand NeoHookean: scenarios.VARIANTS),
loaded by the reference's OWN scene layer (SimContext + ForceBuilder + mclscene, oracle/scene_export.cpp) with the GUI
samples' setup() restated headlessly, exported as fixtures (tests/golden/shipped_*.scene.npz) and run on the device.

 * free-running against the committed reference trajectories: 1e-9 on the reproducible scenes (windyflag: cloth +
   hinges + anchors + wind; plinkopony: ARAP + 23 collision cylinders); the hyperelastic scenes are reported against
   the reference's own sensitivity (bunnyexpand, scrambled, is fully chaotic: the reference differs from itself by
   O(1) after a 1e-15 perturbation);
 * teacher-forced against the unmodified reference run side by side (when oracle/_ref travelled): every iteration's
   local step bit-exact (tets, triangles, hinges, anchors, collisions), global step within 1e-9.
"""
import os

import numpy as np
import pytest

from scenarios import DevAdapter, RefAdapter, build_shipped, run_scenario
from util import GOLDEN, TOL_ITER, have_ref, rel_l2

pytestmark = pytest.mark.gpu
SHIP = build_shipped(GOLDEN)


def _forced_first_frame(sc, gold):
    ad = DevAdapter(sc["scene"])
    sim = ad.sim
    ev = sc.get("events")
    if ev is not None:
        ev(0, ad)
    xi = gold["x_it"][0]
    K = xi.shape[0]
    worst = 0.0
    for k in range(K):
        sim.debug_local_step(xi[k])
        sim.debug_global_step(xi[0])        # x_it[0] is x_bar = x + dt v of the frame (the first iterate)
        x_next = xi[k + 1] if k + 1 < K else gold["x"][0]
        worst = max(worst, rel_l2(sim.x_iter, x_next))
    ad.close()
    return worst


@pytest.mark.parametrize("name", list(SHIP))
def test_shipped_scene_free_running(name):
    gold = np.load(os.path.join(GOLDEN, f"shipped_{name}.ref.npz"))
    sc = SHIP[name]
    ad = DevAdapter(sc["scene"])
    res = run_scenario(ad, sc, dump=False)
    ad.close()
    F = gold["x"].shape[0]
    err = max(rel_l2(res["x"][f], gold["x"][f]) for f in range(F))
    sens = float(gold["sens_x"])
    reproducible = sens < 1e-9
    tol = TOL_ITER if reproducible else 30.0 * sens
    print(f"shipped {name}: {F} frames, max rel-L2 of x {err:.2e}; reference self-sensitivity {sens:.1e}; "
          f"{'gate 1e-9' if reproducible else 'chaotic in the reference itself, gate 30 x sensitivity'}")
    assert np.all(np.isfinite(res["x"]))
    if tol < 1e-2:
        assert err <= tol
    else:
        # The reference run is chaotic (a 1e-15 perturbation of ITS OWN input moves its trajectory by `sens`): a free-running
        # comparison says nothing, so the committed reference iterates of frame 0 are replayed instead -- x entering every
        # ADMM iteration is the reference's, u and the optimiser state are carried by the device -- and x leaving each
        # iteration must match the reference's to the per-iteration gate.  (The live teacher-forced test below additionally
        # checks z, u and the optimiser state bit for bit when the reference library travelled with the snapshot.)
        worst = _forced_first_frame(sc, gold)
        print(f"shipped {name}: frame 0 replayed from the committed reference iterates: worst rel-L2 of x per iteration {worst:.2e}")
        assert worst <= TOL_ITER


@pytest.mark.parametrize("name", list(SHIP))
def test_shipped_scene_teacher_forced_live(name):
    if not have_ref():
        pytest.skip("oracle/_ref/libadmm_ref.so not present on this box")
    sc = dict(SHIP[name])
    sc["frames"] = min(sc["frames"], 6 if not name.startswith("poordillo") else (22 if name == "poordillo" else 8))   # poordillo: include the release at frame 20
    ra = RefAdapter(sc["scene"])
    gold = run_scenario(ra, sc, dump=True)
    ra.close()
    ad = DevAdapter(sc["scene"])
    sim = ad.sim
    F, K = gold["x_it"].shape[:2]
    R = gold["z_it"].shape[2]
    has_prox = "prox_it" in gold and gold["prox_it"].size > 0
    u_prev = np.zeros(R)
    prox_prev = np.ones(gold["prox_it"].shape[2:]) if has_prox else None
    ev = sc.get("events")
    n_exact = n_total = 0
    worst_local = worst_x = 0.0
    for f in range(F):
        if ev is not None:
            ev(f, ad)
        for k in range(K):
            sim.set_state(2, u_prev)
            if has_prox:
                sim.set_state(3, prox_prev)
            sim.debug_local_step(gold["x_it"][f, k])
            z, u = sim.z, sim.u
            gz, gu = gold["z_it"][f, k], gold["u_it"][f, k]
            n_total += 2
            n_exact += int(np.array_equal(z, gz)) + int(np.array_equal(u, gu))
            worst_local = max(worst_local, float(np.abs(z - gz).max()), float(np.abs(u - gu).max()))
            sim.debug_global_step(gold["x_it"][f, 0])
            x_next = gold["x_it"][f, k + 1] if k + 1 < K else gold["x"][f]
            worst_x = max(worst_x, rel_l2(sim.x_iter, x_next))
            u_prev = gu
            if has_prox:
                prox_prev = gold["prox_it"][f, k]
    ad.close()
    print(f"shipped {name} teacher-forced: {F} frames x {K} iterations, local step bit-exact in {n_exact}/{n_total} vectors "
          f"(worst abs {worst_local:.1e}), global step worst rel-L2 {worst_x:.1e}")
    assert worst_x <= TOL_ITER
    assert n_exact == n_total
