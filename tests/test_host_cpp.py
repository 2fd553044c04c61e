"""The C++ host layer (admm-elastic-sca_b200/host: namespace admm on top of the C ABI).

 * the reference's OWN console samples (A/samples/singletet.cpp, singlenode.cpp), compiled unmodified against
   host/*.hpp (host/Makefile), must print the reference's known answers: 171.571 and
   -9.8, -29.4, -58.8, -98;
 * host_check builds golden scenes the way ForceBuilder does (one Force object per element) and must reproduce
   the C-ABI path bit for bit and the golden reference within the free-running gate.
"""
import os
import re
import subprocess

import numpy as np
import pytest

import scenes
from scenarios import DevAdapter, build_scenarios, run_scenario
from util import GOLDEN, ROOT, TOL_ITER, rel_l2

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "admm-elastic-sca_b200", "host", "build")


def _need(binary):
    p = os.path.join(HOST, binary)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (host/Makefile needs Eigen from the reference tree)")
    return p


def test_reference_singletet_sample_runs_on_the_gpu_solver():
    out = subprocess.run([_need("ref_singletet")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    m = re.search(r"Node 4 x: ([-0-9.e+]+)", out.stdout)
    assert m, out.stdout
    assert m.group(1) == "171.571"   # A/samples/singletet.cpp:49 prints with default precision


def test_plain_c_example_of_the_abi():
    """examples/abi_minimal.c (C99, -pedantic -Werror): the singletet scene straight through the C ABI."""
    out = subprocess.run([_need("abi_minimal")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "Node 4 x: 171.571"


def test_reference_singlenode_sample_runs_on_the_gpu_solver():
    out = subprocess.run([_need("ref_singlenode")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    ys = [float(v) for v in re.findall(r"pos: \(0, ([-0-9.e+]+), 0\)", out.stdout)]
    assert ys == [-9.8, -29.4, -58.8, -98.0], out.stdout


def write_scene_txt(path, sc, frames):
    x = np.asarray(sc["x"], dtype=np.float64).reshape(-1, 3)
    L = [f"{float(sc['dt'])!r} {int(sc['iters'])} {frames} {x.shape[0]}"]
    for i in range(x.shape[0]):
        L.append(f"{float(x[i, 0])!r} {float(x[i, 1])!r} {float(x[i, 2])!r} {float(sc['m'][i])!r}")
    if "x_after_init" in sc:
        L.append("1")
        L.append(" ".join(repr(float(v)) for v in np.asarray(sc["x_after_init"]).reshape(-1)))
    else:
        L.append("0")
    coll = [b for b in sc["batches"] if b["type"] == "collision"]
    rest = [b for b in sc["batches"] if b["type"] != "collision"]
    assert not coll or sc["batches"][-1]["type"] == "collision"
    L.append(str(len(rest)))
    for b in rest:
        t = b["type"]
        idx = np.asarray(b["idx"])
        if t == "tets":
            L.append(f"tets {b['kind']} {len(idx)} {float(b.get('p0', 0))!r} {float(b.get('p1', 0))!r} {float(b.get('p2', 0))!r} {int(b.get('maxit', 10))} 0")
            L += [" ".join(map(str, r)) for r in idx]
        elif t == "tris":
            L.append(f"tris {b['kind']} {len(idx)} {float(b['stiffness'])!r} {float(b.get('lmin', 0.0))!r} {float(b.get('lmax', 9999999.0))!r} 0 {int(b.get('flag', 1))}")
            L += [" ".join(map(str, r)) for r in idx]
        elif t == "springs":
            k = np.broadcast_to(b["stiffness"], (len(idx),))
            L.append(f"springs 0 {len(idx)} 0 0 0 0 0")
            L += [f"{r[0]} {r[1]} {float(ki)!r}" for r, ki in zip(idx, k)]
        elif t == "bends":
            L.append(f"bends 0 {len(idx)} {float(b['stiffness'])!r} 0 0 0 0")
            L += [" ".join(map(str, r)) for r in idx]
        elif t == "static_anchors":
            L.append(f"static_anchors 0 {idx.size} {float(b.get('weight', -1.0))!r} 0 0 0 0")
            L += [str(int(i)) for i in idx.reshape(-1)]
        elif t == "moving_anchors":
            L.append(f"moving_anchors 0 {idx.size} {float(b.get('weight', -1.0))!r} 0 0 0 0")
            L += [f"{int(i)} {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}" for i, p in zip(idx.reshape(-1), np.asarray(b["pos"], dtype=np.float64))]
    if coll:
        c = coll[0]
        L.append(str(len(c["kinds"])))
        L.append(repr(float(c.get("weight", 32.0))))
        for k, p in zip(c["kinds"], np.asarray(c["params"], dtype=np.float64)):
            L.append(f"{int(k)} {float(p[0])!r} {float(p[1])!r} {float(p[2])!r} {float(p[3])!r}")
    else:
        L.append("0")
    ex = sc.get("explicit", [])
    L.append(str(len(ex)))
    for e in ex:
        d = np.asarray(e["dir"], dtype=np.float64)
        if e["type"] == "gravity":
            L.append(f"gravity {float(d[0])!r} {float(d[1])!r} {float(d[2])!r}")
        else:
            tr = np.asarray(e["tris"])
            L.append(f"wind {float(d[0])!r} {float(d[1])!r} {float(d[2])!r}")
            L.append(str(len(tr)))
            L.append(" ".join(map(str, tr.reshape(-1))))
    open(path, "w").write("\n".join(L) + "\n")


SCEN = build_scenarios()
NAMES = [n for n in SCEN if "events" not in SCEN[n] and n != "singlenode"]


@pytest.mark.parametrize("name", NAMES)
def test_cpp_host_layer_matches_c_abi_and_golden(name, tmp_path):
    binary = _need("host_check")
    gold = np.load(os.path.join(GOLDEN, f"{name}.ref.npz"))
    sc = scenes.load_scene(os.path.join(GOLDEN, f"{name}.scene.npz"))
    frames = SCEN[name]["frames"]
    txt, out = str(tmp_path / "scene.txt"), str(tmp_path / "x.bin")
    write_scene_txt(txt, sc, frames)
    r = subprocess.run([binary, txt, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    xs = np.fromfile(out, dtype=np.float64).reshape(frames, -1)
    ad = DevAdapter(sc)
    res = run_scenario(ad, dict(scene=sc, frames=frames), dump=False)
    ad.close()
    same = np.array_equal(xs, res["x"])
    err = max(rel_l2(xs[f], gold["x"][f]) for f in range(frames))
    tol = max(TOL_ITER, 30.0 * float(gold["sens_x"]))
    print(f"{name}: C++ host layer vs C ABI bit-identical: {same}; vs golden x {err:.1e} (gate {tol:.1e})")
    assert err <= tol
    assert max(rel_l2(xs[f], res["x"][f]) for f in range(frames)) <= tol


def test_user_defined_explicit_force_keeps_all_explicit_forces_on_the_host(tmp_path):
    """The host layer runs the reference's own ExplicitForce / WindForce on the device; a user subclass in the list sends
    EVERY explicit force through its host project() in list order instead (no reordering, no double application).  Wind
    reads the velocities gravity has just changed, so the order matters: both routes must give the same frames."""
    binary = _need("host_check")
    sc = scenes.load_scene(os.path.join(GOLDEN, "cloth6x4.scene.npz"))
    assert [e["type"] for e in sc["explicit"]] == ["gravity", "wind"]
    frames = 3
    txt = str(tmp_path / "scene.txt")
    write_scene_txt(txt, sc, frames)
    xs = []
    for mode in ([], ["userforce"]):
        out = str(tmp_path / f"x{len(mode)}.bin")
        r = subprocess.run([binary, txt, out] + mode, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        xs.append(np.fromfile(out, dtype=np.float64).reshape(frames, -1))
    err = max(rel_l2(xs[1][f], xs[0][f]) for f in range(frames))
    print(f"device explicit forces vs host explicit forces (user subclass present): rel-L2 {err:.1e}")
    assert err <= 1e-10   # the solve accumulates with atomics (last bits vary) and the wind drag amplifies that ~3x per frame


@pytest.mark.parametrize("name", NAMES)
def test_soa_force_batches_are_bit_identical_to_one_object_per_force(name, tmp_path):
    """a run of elements registered as ONE admm::TetBatch / TriangleBatch / BendBatch / SpringBatch object
    (what host/scene/ForceBuilderBatched.cpp builds) must give the same frames, bit for bit, as the reference's one heap
    object per element; the per-element weights are read back into the batch and survive recompute_weights()."""
    binary = _need("host_check")
    sc = scenes.load_scene(os.path.join(GOLDEN, f"{name}.scene.npz"))
    frames = SCEN[name]["frames"]
    txt = str(tmp_path / "scene.txt")
    write_scene_txt(txt, sc, frames)
    env = dict(os.environ, ADMMB_DETERMINISTIC="1")
    xs, notes = [], []
    for mode in ([], ["soa"]):
        out = str(tmp_path / f"x{len(mode)}.bin")
        r = subprocess.run([binary, txt, out] + mode, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        xs.append(np.fromfile(out, dtype=np.float64).reshape(frames, -1))
        notes.append(r.stdout.strip().splitlines()[-1])
    print(f"{name}: {notes[0]} | {notes[1]}")
    assert np.array_equal(xs[0], xs[1])
