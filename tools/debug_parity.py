"""Developer aid: per-(frame, iteration) error table of one scenario against its golden dump."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "admm-elastic-sca_b200", "pyhost"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import scenes
from scenarios import DevAdapter, build_scenarios, run_scenario
from util import GOLDEN, rel_l2

name = sys.argv[1]; solver = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = build_scenarios()
gold = np.load(os.path.join(GOLDEN, f"{name}.ref.npz"))
sc = dict(S[name]); sc["scene"] = scenes.load_scene(os.path.join(GOLDEN, f"{name}.scene.npz"))
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-12
ad = DevAdapter(sc["scene"], solver=solver, cg_tol=tol)
res = run_scenario(ad, sc, dump=True)
F, K = res["x_it"].shape[:2]
for f in range(F):
    for k in range(K):
        print(f"f{f} it{k:2d}  x {rel_l2(res['x_it'][f,k], gold['x_it'][f,k]):.2e}  z {rel_l2(res['z_it'][f,k], gold['z_it'][f,k]):.2e}  u {rel_l2(res['u_it'][f,k], gold['u_it'][f,k]):.2e}")
    print(f"f{f} final x {rel_l2(res['x'][f], gold['x'][f]):.2e} v {rel_l2(res['v'][f], gold['v'][f]):.2e}")
    # which rows differ most in z at the first bad iteration
for f in range(F):
    for k in range(K):
        d = np.abs(res['z_it'][f,k] - gold['z_it'][f,k])
        if d.max() > 1e-9:
            rows = np.argsort(-d)[:12]
            print("first bad z at", f, k, "rows", rows, "forces", rows // 9, d[rows])
            break
    else:
        continue
    break
if "prox_iters" in gold.files:
    print("its equal frac", np.mean(res["prox_iters"] == gold["prox_iters"]), "state err", np.abs(res["prox_state"] - gold["prox_state"]).max())
print(ad.sim.info())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"dev_{name}.npz"), **res)
