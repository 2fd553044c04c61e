"""Lane-divergence diagnostics of the hyperelastic local step: per-tet line-search trial counts (ADMMB_STATE_PROX_TRIALS) of
consecutive ADMM iterations in the conditioned regime -- how far a warp's maximum is above its mean (idle lanes), how well a
tet's count predicts its next one, and what sorting a block's tets by their previous count would save."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "admm-elastic-sca_b200", "pyhost"), os.path.join(ROOT, "tests")]
import numpy as np
import admm_b200, scenes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sc = scenes.cube_scene(N, kind=scenes.TET_NH, mu=1e5, lam=1e5, maxit=5, mass=1000.0, dt=0.04, iters=10, stretch=1.3)
sim = admm_b200.System(sc)
sim.set_x(sc["x_after_init"]); sim.upload()
sim.step_resident(frames=25)
sim.step_resident(frames=20)                       # the frames bench.py times (10 iterations each)
h = np.bincount(sim.state(5).astype(np.int32), minlength=22)
print("bench regime, last iteration of frame 45: trial-count histogram", h, " share with 2 trials %.3f, with 20 %.3f, mean %.2f" % (h[2] / h.sum(), h[20] / h.sum(), (h * np.arange(len(h))).sum() / h.sum()))
tr = []
for k in range(4):
    sim.step_resident(frames=1, iters=1)          # one ADMM iteration per call: consecutive local steps
    tr.append(sim.state(5).astype(np.int32))
info = sim.info()
tr = np.array(tr)
T = tr.shape[1]
print("tets", T, "mean trials", tr.mean(axis=1), "hist of last:", np.bincount(tr[-1], minlength=22))
# the device order is the batch's Morton order, not the user order: approximate warps by Morton-sorting here is not possible
# from outside, so use the export order only for the temporal statistics; block / warp statistics use chunks of the export
# order of a Kuhn cube (6 tets per cell, cells x-fastest), which is spatially coherent too
for W in (32, 320):
    n = (T // W) * W
    a = tr[-1][:n].reshape(-1, W)
    print(f"chunks of {W}: mean of max {a.max(axis=1).mean():.2f}, mean {a.mean():.2f}, idle fraction {1 - a.mean() / a.max(axis=1).mean():.3f}")
prev, cur = tr[-2], tr[-1]
print("corr(prev, cur) =", np.corrcoef(prev, cur)[0, 1], " P(same) =", (prev == cur).mean(), " P(|d|<=2) =", (np.abs(prev - cur) <= 2).mean())
# sorting each block of 320 by the PREVIOUS count, warps of 32 inside the block
n = (T // 320) * 320
p = prev[:n].reshape(-1, 320); c = cur[:n].reshape(-1, 320)
order = np.argsort(p, axis=1, kind="stable")
cs = np.take_along_axis(c, order, axis=1).reshape(-1, 32)
cu = c.reshape(-1, 32)
print(f"warp max summed: unsorted {cu.max(axis=1).sum()}, sorted by previous count {cs.max(axis=1).sum()} -> {cs.max(axis=1).sum() / cu.max(axis=1).sum():.3f}; ideal (sorted by current) {np.sort(c, axis=1).reshape(-1, 32).max(axis=1).sum() / cu.max(axis=1).sum():.3f}")
sim.close()
