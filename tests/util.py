"""Shared helpers of the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Parity gates of BASELINE.json north_star / SURVEY.md 8(d): per ADMM iteration x, z, u within 1e-9
# relative L2 (FP64); trajectories within 1e-6.
TOL_ITER = 1e-9
TOL_TRAJ = 1e-6


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    if n == 0.0:
        return d
    return d / n


def have_ref():
    from oracle import ref
    return ref.available()


def apply_host_explicit(system_like, scene, x, v, dt):
    """Per-frame explicit forces that are not device-resident (wind): applied to v by the caller."""
    from admm_b200 import wind_project
    for e in scene.get("explicit", []):
        if e["type"] == "wind":
            wind_project(x, v, e["tris"], e["dir"], dt)
