"""Pins the plain-C oracle restatement (oracle/port/admm_oracle.c) against the reference's golden dumps
(tests/golden/*.npz, generated from the unmodified reference) -- and, where oracle/_ref is present, against the
reference's two known answers directly.  CPU only."""
import os

import numpy as np
import pytest

import scenes
from scenarios import build_scenarios, run_scenario
from util import GOLDEN, TOL_ITER, rel_l2

SCEN = build_scenarios()


def _port():
    from oracle import port
    if not port.available():
        pytest.skip("oracle/libadmm_oracle.so not built (run __graft_entry__.build())")
    return port


def _worst(a, g):
    if a.size == 0:
        return 0.0
    if a.ndim == 3:
        return max(rel_l2(a[f, k], g[f, k]) for f in range(a.shape[0]) for k in range(a.shape[1]))
    return max(rel_l2(a[f], g[f]) for f in range(a.shape[0]))


def test_known_answers_singletet_and_singlenode():
    """the only known-answer programs of the reference."""
    port = _port()
    sc = scenes.singletet_scene()
    ad = port.PortAdapter(sc)
    ad.set_x(sc["x_after_init"])
    x, _ = ad.step()
    assert "%g" % x[9] == "171.571"           # A/samples/singletet.cpp:49
    ad.close()
    ad = port.PortAdapter(scenes.singlenode_scene())
    ys = ["%g" % ad.step()[0][1] for _ in range(4)]
    assert ys == ["-9.8", "-29.4", "-58.8", "-98"]   # A/samples/singlenode.cpp:44
    ad.close()


@pytest.mark.parametrize("name", list(SCEN))
def test_port_against_reference_golden(name):
    port = _port()
    gold = np.load(os.path.join(GOLDEN, f"{name}.ref.npz"))
    scenario = dict(SCEN[name])
    scenario["scene"] = scenes.load_scene(os.path.join(GOLDEN, f"{name}.scene.npz"))
    ad = port.PortAdapter(scenario["scene"])
    res = run_scenario(ad, scenario, dump=True)
    ad.close()
    report, ok = [], True
    for key in ("x_it", "z_it", "u_it", "x", "v"):
        err, sens = _worst(res[key], gold[key]), float(gold["sens_" + key])
        tol = max(TOL_ITER, 30.0 * sens)
        report.append(f"{key} {err:.1e} (gate {tol:.1e})")
        ok = ok and err <= tol
    print(f"{name}: " + "; ".join(report))
    assert ok, "; ".join(report)


@pytest.mark.parametrize("name", list(SCEN))
def test_port_local_step_is_bit_exact_teacher_forced(name):
    """The restatement is literal: replayed from the reference's own inputs (x entering the iteration, u and optimiser
    state of the previous one), every ADMM iteration's z, u and optimiser state come out bit for bit."""
    port = _port()
    gold = np.load(os.path.join(GOLDEN, f"{name}.ref.npz"))
    scenario = dict(SCEN[name])
    scenario["scene"] = scenes.load_scene(os.path.join(GOLDEN, f"{name}.scene.npz"))
    ad = port.PortAdapter(scenario["scene"])
    F, K = gold["x_it"].shape[:2]
    R = gold["z_it"].shape[2]
    has_prox = "prox_it" in gold.files and gold["prox_it"].size > 0
    u_prev = np.zeros(R)
    prox_prev = np.ones(gold["prox_it"].shape[2:]) if has_prox else None
    ev = scenario.get("events")
    n_exact = n_total = 0
    for f in range(F):
        if ev is not None:
            ev(f, ad)
        for k in range(K):
            z, u, p = ad.local_step(gold["x_it"][f, k], u_prev, prox_prev)
            ok = np.array_equal(z, gold["z_it"][f, k]) and np.array_equal(u, gold["u_it"][f, k])
            if has_prox:
                ok = ok and np.array_equal(p, gold["prox_it"][f, k])
                prox_prev = gold["prox_it"][f, k]
            n_total += 1
            n_exact += int(ok)
            u_prev = gold["u_it"][f, k]
    ad.close()
    print(f"{name}: port local step bit-exact in {n_exact}/{n_total} iterations")
    assert n_exact == n_total


def test_reference_library_known_answers_when_present():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built on this box")
    sc = scenes.singletet_scene()
    r = ref.RefSystem(sc)
    r.set_x(sc["x_after_init"])
    r.step()
    assert "%g" % r.x[9] == "171.571"
