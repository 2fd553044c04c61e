"""The device cannot call the host libm, and CUDA's log / exp differ from glibc's in the last bit for a few percent of
arguments -- enough to flip branches of the reference's truncated line search.  csrc/elastic_math.h therefore carries
clones of glibc 2.39's `__log_fma` / `__exp_fma` (tables read from this image's libm.so.6 by tools/extract_glibc_*.py).
Here the clones, compiled for the host by tests/hostcheck, are compared with the libm of this box bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

HC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck", "libhostcheck.so")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _lib():
    if not os.path.exists(HC):
        pytest.skip("tests/hostcheck not built (run __graft_entry__.build())")
    L = C.CDLL(HC)
    for f in (L.hc_exp_mismatches, L.hc_log_mismatches):
        f.restype = C.c_long
        f.argtypes = [C.c_long, _dp]
    return L


def test_exp_clone_matches_libm_bit_for_bit():
    L = _lib()
    rng = np.random.default_rng(1)
    sets = [rng.uniform(-20, 20, 2_000_000), rng.uniform(-1, 1, 1_000_000), rng.uniform(-1e-10, 1e-10, 200_000),
            rng.uniform(600, 720, 300_000), rng.uniform(-760, -600, 300_000), rng.uniform(-1100, 1100, 300_000),
            np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 709.782712893384, 709.8, -745.2, -745.13321910194122, -708.4, 1e-300,
                      -1e-300, 5e-324, 1024.0, -1075.0, 2.0 ** -54, 2.0 ** -55])]
    for x in sets:
        assert L.hc_exp_mismatches(len(x), np.ascontiguousarray(x)) == 0


def test_log_clone_matches_libm_bit_for_bit():
    L = _lib()
    rng = np.random.default_rng(2)
    sets = [np.exp(rng.uniform(-30, 30, 2_000_000)), rng.uniform(0.9, 1.1, 1_000_000), rng.uniform(1e-310, 1e-300, 100_000),
            np.array([0.0, -0.0, 1.0, np.inf, -1.0, np.nan, 5e-324, 1.7976931348623157e308])]
    for x in sets:
        assert L.hc_log_mismatches(len(x), np.ascontiguousarray(x)) == 0


@pytest.mark.parametrize("tool,inc", [("extract_glibc_exp.py", "glibc_exp_data.inc"), ("extract_glibc_log.py", "glibc_log_data.inc")])
def test_committed_table_is_what_this_libm_holds(tool, inc):
    """Provenance of the constants: the committed table equals what tools/extract_glibc_exp.py reads from the libm of this
    box (skipped if libm moved its tables, in which case the sweep tests above are the authority)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", tool)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("libm layout differs from glibc 2.39-0ubuntu8.5: " + r.stderr.strip().splitlines()[-1])
    committed = open(os.path.join(root, "admm-elastic-sca_b200", "csrc", inc)).read()
    import re
    hexes = lambda t: [int(h, 16) for h in re.findall(r"0x[0-9a-fA-F]{16}", t)]
    assert hexes(r.stdout) == hexes(committed) and len(hexes(committed)) > 200
