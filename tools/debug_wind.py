import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "admm-elastic-sca_b200", "pyhost"))
import numpy as np
import admm_b200, scenes
sc = scenes.cloth_scene(9, 7, springs=True, wind=(10.0, 0.0, 2.0), iters=10, name="cloth_wind")
print([e["type"] for e in sc["explicit"]])
for iters in (0, 10):
    a = admm_b200.System(sc, iters=iters)
    b = admm_b200.System(sc, iters=iters, host_explicit=True)
    for f in range(4):
        b.apply_host_explicit()
        a.step(); b.step()
        dx = np.abs(a.m_x - b.m_x); dv = np.abs(a.m_v - b.m_v)
        print("iters", iters, "frame", f, "x mismatches", int((dx > 0).sum()), "max", dx.max(), "v mismatches", int((dv > 0).sum()), dv.max(), "first", np.nonzero(dx > 0)[0][:6])
    a.close(); b.close()
