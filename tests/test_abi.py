"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to run without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import admm_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exists_and_exports_every_declared_symbol():
    assert os.path.exists(admm_b200.LIB_PATH), "build with __graft_entry__.build()"
    L = C.CDLL(admm_b200.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "admm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(admmb_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found in the header"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/admm_b200.h but not exported"
    assert sorted(admm_b200.EXPORTS) == declared


def test_version_string():
    L = admm_b200.lib()
    assert b"sm_100a" in L.admmb_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; the failure path is exercised on the CPU-only box")
    L = admm_b200.lib()
    h = C.c_void_p()
    rc = L.admmb_create(0, C.byref(h))
    assert rc == -3  # ADMMB_E_CUDA
    assert b"no CPU path" in L.admmb_last_error(None)


def test_every_entry_point_is_documented_and_cites_the_reference():
    """INTEGRATION.md maps every exported entry point to the reference interface it replaces; the header cites
    reference file:line for the path's entry points."""
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [name for name in admm_b200.EXPORTS if name not in integ]
    assert not missing, f"not in INTEGRATION.md: {missing}"
    header = open(os.path.join(ROOT, "include", "admm_b200.h")).read()
    cites = re.findall(r"[A-Za-z]+\.(?:cpp|hpp|h):\d+", header)
    assert len(cites) >= 20, "the header should cite the reference (file:line) for the entry points it replaces"
