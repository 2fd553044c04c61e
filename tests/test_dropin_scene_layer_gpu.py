"""Drop-in at the level of the reference's scene layer: the reference's OWN SimContext + ForceBuilder + mclscene (XML
loader, tet / triangle meshes), compiled unmodified against admm-elastic-sca_b200/host (tests/dropin/Makefile), load the
four shipped sample scenes from their XML files and step them on the B200 through SimContext::step().  The positions
after every frame are compared with the goldens the unmodified reference solver produced for the same scenes
(tests/golden/shipped_*.ref.npz) under the free-running gate: 1e-9 where the reference is reproducible (windyflag,
plinkopony), 30 x its own sensitivity where it is not (bunnyexpand, poordillo: DESIGN.md section 5)."""
import os
import subprocess

import numpy as np
import pytest

from scenarios import SHIPPED_FRAMES
from util import GOLDEN, ROOT, TOL_ITER, rel_l2

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(ROOT, "tests", "dropin", "build")
XML = dict(bunnyexpand="bunnyexpand.xml", windyflag="cloth.xml", poordillo="poordillo.xml", plinkopony="plinko.xml")


@pytest.mark.parametrize("name", list(XML))
def test_reference_scene_layer_runs_on_the_gpu_solver(name, tmp_path):
    runner = os.path.join(DROPIN, "ref_scene_runner")
    archive = os.path.join(DROPIN, "scenes.tar")
    if not (os.path.exists(runner) and os.path.exists(archive)):
        pytest.skip("tests/dropin not built (needs the reference tree at build time: __graft_entry__.build())")
    import tarfile
    with tarfile.open(archive) as tf:
        tf.extractall(tmp_path, members=[m for m in tf.getmembers() if m.name.startswith(name + "/")], filter="data")
    xml = str(tmp_path / name / XML[name])
    gold = np.load(os.path.join(GOLDEN, f"shipped_{name}.ref.npz"))
    frames = SHIPPED_FRAMES[name]
    out = str(tmp_path / "x.bin")
    r = subprocess.run([runner, name, xml, str(frames), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    xs = np.fromfile(out, dtype=np.float64).reshape(frames, -1)
    assert xs.shape == gold["x"].shape
    err = max(rel_l2(xs[f], gold["x"][f]) for f in range(frames))
    tol = max(TOL_ITER, 30.0 * float(gold["sens_x"]))
    print(f"{name}: {r.stdout.strip().splitlines()[-1]}; x vs the reference's golden trajectory {err:.1e} (gate {tol:.1e})")
    assert np.isfinite(xs).all()
    assert err <= tol
    if tol > 1e-2 or name == "windyflag":
        # Where the reference itself is chaotic the gate above is loose (bunnyexpand: vacuous).  The sharp statement is that
        # the reference's scene layer on top of the host layer feeds the solver EXACTLY what the exported fixture feeds it
        # through the C ABI: with the bit-reproducible solve on both sides the two trajectories must be identical in every
        # bit, and the fixture-driven run is the one the teacher-forced tests pin to the reference (tests/test_shipped_scenes_gpu.py).
        from scenarios import DevAdapter, build_shipped, run_scenario
        sc = build_shipped(GOLDEN, variants=False)[name]
        env = dict(os.environ, ADMMB_DETERMINISTIC="1")
        out2 = str(tmp_path / "x_det.bin")
        r2 = subprocess.run([runner, name, xml, str(frames), out2], capture_output=True, text=True, timeout=600, env=env)
        assert r2.returncode == 0, r2.stdout + r2.stderr
        xd = np.fromfile(out2, dtype=np.float64).reshape(frames, -1)
        ad = DevAdapter(sc["scene"], deterministic=True)
        res = run_scenario(ad, sc, dump=False)
        ad.close()
        same = np.array_equal(xd, res["x"])
        print(f"{name}: scene layer + host layer vs fixture through the C ABI, deterministic solve: bit-identical = {same}")
        assert same


@pytest.mark.parametrize("name", list(XML))
def test_batched_force_factory_gives_the_same_trajectory(name, tmp_path):
    """the scene layer with host/scene/ForceBuilderBatched.cpp swapped in for src/ForceBuilder.cpp (one SoA
    batch per object and force instead of one heap object per element, hashed hinge dedupe) steps the shipped scenes to the
    same positions, bit for bit, as the scene layer with the reference's own factory (deterministic solve on both sides)."""
    runners = [os.path.join(DROPIN, "ref_scene_runner"), os.path.join(DROPIN, "ref_scene_runner_batched")]
    archive = os.path.join(DROPIN, "scenes.tar")
    if not (all(os.path.exists(r) for r in runners) and os.path.exists(archive)):
        pytest.skip("tests/dropin not built (needs the reference tree at build time: __graft_entry__.build())")
    import tarfile
    with tarfile.open(archive) as tf:
        tf.extractall(tmp_path, members=[m for m in tf.getmembers() if m.name.startswith(name + "/")], filter="data")
    xml = str(tmp_path / name / XML[name])
    frames = SHIPPED_FRAMES[name] if name == "poordillo" else min(SHIPPED_FRAMES[name], 12)   # poordillo: through the release + recompute_weights() at frame 20
    env = dict(os.environ, ADMMB_DETERMINISTIC="1")
    xs = []
    for k, runner in enumerate(runners):
        out = str(tmp_path / f"x{k}.bin")
        r = subprocess.run([runner, name, xml, str(frames), out], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        xs.append(np.fromfile(out, dtype=np.float64).reshape(frames, -1))
    same = np.array_equal(xs[0], xs[1])
    print(f"{name}: {frames} frames, batched factory vs the reference's factory: bit-identical = {same}")
    assert np.isfinite(xs[1]).all()
    assert same
