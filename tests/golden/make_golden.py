"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/libadmm_ref.so, built by oracle/Makefile from /root/reference) on the parity scenarios.

Run in the build container (needs /root/reference for the _ref build and the shipped meshes):
    python tests/golden/make_golden.py
Each <name>.npz holds the scene (inputs) and the reference's per-iteration x / z / u dumps, per-frame x / v
and the hyperelastic optimiser state; the -m gpu tests and the oracle-port tests compare against them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "admm-elastic-sca_b200", "pyhost"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import scenes  # noqa: E402
from scenarios import RefAdapter, build_scenarios, run_scenario  # noqa: E402


# The reference's own sensitivity: the same run from initial positions perturbed by 1e-15, for SENS_SEEDS different
# sign patterns; the stored value is the MAXIMUM over the seeds.  The response is heavy-tailed (the truncated line search
# bifurcates: e.g. cube2_stvk answers 11 of 12 perturbations with 5e-8 and one with 6e-6), so a single sample
# under-states what a 1-ulp difference can do.
SENS_SEEDS = tuple(range(99, 99 + 16))


def sensitivity(sc, res, keys):
    out = {"sens_" + k: 0.0 for k in keys}
    for seed in SENS_SEEDS:
        ad = RefAdapter(sc["scene"])
        per = run_scenario(ad, sc, dump=True, perturb=1e-15, seed=seed)
        ad.close()
        for key in keys:
            a, g = per[key], res[key]
            if a.size == 0:
                continue
            flat = a.reshape(-1, a.shape[-1]), g.reshape(-1, g.shape[-1])
            num = np.linalg.norm(flat[0] - flat[1], axis=1)
            den = np.linalg.norm(flat[1], axis=1)
            out["sens_" + key] = max(out["sens_" + key], float(np.max(np.where(den > 0, num / np.where(den > 0, den, 1.0), num))))
    return {k: np.float64(v) for k, v in out.items()}


def main():
    S = build_scenarios()
    for name, sc in S.items():
        ad = RefAdapter(sc["scene"])
        res = run_scenario(ad, sc, dump=True)
        ad.close()
        res.update(sensitivity(sc, res, ("x_it", "z_it", "u_it", "x", "v")))
        scenes.save_scene(os.path.join(HERE, f"{name}.scene.npz"), sc["scene"])
        np.savez_compressed(os.path.join(HERE, f"{name}.ref.npz"), **res)
        sz = os.path.getsize(os.path.join(HERE, f"{name}.ref.npz")) / 1024
        print(f"{name:14s} frames={sc['frames']} x_it{res['x_it'].shape} z_it{res['z_it'].shape} -> {sz:.0f} KiB   "
              f"self-sensitivity x_it {res['sens_x_it']:.1e} z_it {res['sens_z_it']:.1e} u_it {res['sens_u_it']:.1e} v {res['sens_v']:.1e}")


def shipped(only=None, export=True):
    """The reference's four shipped scenes, loaded by the reference's own scene layer (oracle/_ref/scene_export), and
    their material variants (scenarios.VARIANTS).  only = names to (re)generate."""
    import subprocess
    import tempfile
    from scenarios import SHIPPED, SHIPPED_FRAMES, build_shipped, parse_exported_scene
    exe = os.path.join(ROOT, "oracle", "_ref", "scene_export")
    if not os.path.exists(exe):
        print("oracle/_ref/scene_export not built (make -C oracle scene_export): skipping the shipped scenes")
        return
    tmp = tempfile.mkdtemp()
    for name in (SHIPPED if export else ()):
        txt = os.path.join(tmp, name + ".txt")
        subprocess.run([exe, name, txt], check=True, stdout=subprocess.DEVNULL)
        scenes.save_scene(os.path.join(HERE, f"shipped_{name}.scene.npz"), parse_exported_scene(txt, name))
    for name, sc in build_shipped(HERE).items():
        if only and name not in only:
            continue
        ad = RefAdapter(sc["scene"])
        res = run_scenario(ad, sc, dump=True)
        ad.close()
        out = dict(x_it=res["x_it"][:1], x=res["x"], v=res["v"][-1:])   # z/u dumps of these scenes are too large to commit
        out.update(sensitivity(sc, res, ("x_it", "x")))
        np.savez_compressed(os.path.join(HERE, f"shipped_{name}.ref.npz"), **out)
        print(f"shipped {name:12s} frames={sc['frames']} nodes={sc['scene']['x'].shape[0]} rows={res['z_it'].shape[2]} "
              f"self-sensitivity x_it {out['sens_x_it']:.1e} x {out['sens_x']:.1e}")


def long_runs():
    """100-frame reference trajectories of the reproducible scenes (every 10th frame kept) + the reference's own sensitivity
    over the same horizon (4 perturbation patterns: each is a 100-frame reference run)."""
    from scenarios import build_long
    for name, sc in build_long(HERE).items():
        ad = RefAdapter(sc["scene"])
        res = run_scenario(ad, sc, dump=False)
        ad.close()
        keep = list(range(9, sc["frames"], 10))
        sens = 0.0
        for seed in SENS_SEEDS[:4]:
            ad = RefAdapter(sc["scene"])
            per = run_scenario(ad, sc, dump=False, perturb=1e-15, seed=seed)
            ad.close()
            sens = max(sens, max(float(np.linalg.norm(per["x"][f] - res["x"][f]) / np.linalg.norm(res["x"][f])) for f in keep))
        np.savez_compressed(os.path.join(HERE, f"long_{name}.ref.npz"), x=res["x"][keep], frames=np.array(keep), sens_x=np.float64(sens))
        print(f"long {name:26s} frames={sc['frames']} kept={len(keep)} self-sensitivity x {sens:.1e}")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        main()
        shipped()
        long_runs()
    elif what == "shipped":
        shipped()
    elif what == "variants":     # only the material variants: the scene fixtures of the four shipped scenes stay as they are
        from scenarios import VARIANTS
        shipped(only=set(VARIANTS), export=False)
    elif what == "long":
        long_runs()
