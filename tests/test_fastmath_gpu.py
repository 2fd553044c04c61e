"""The local step's exact fast paths for division / reciprocal / square root (csrc/elastic_math.h: recip_of + div_by,
rcp_x, sqrt_x, each with its shared fallback to the operator) against the plain IEEE operators on the device.

These helpers issue the very instruction chain nvcc emits for `a / y`, `1.0 / x` and `sqrt(x)` -- they only move the
range test to the caller so that several operations share one fallback branch and divisions by the same denominator
share the reciprocal refinement.  The reference (x86-64) computes correctly rounded quotients / roots; so does the
operator on the device; so must the fast paths, in EVERY bit, for every operand class: any bit pattern (NaN, inf,
subnormals), moderate magnitudes, values near 1, extreme exponents, zeros and powers of two, equal operands.
"""
import numpy as np
import pytest

import admm_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 20261017])
def test_fast_division_reciprocal_sqrt_are_bit_exact(seed):
    mism, fallbacks = admm_b200.fastmath_selftest(samples=1 << 27, seed=seed)
    assert mism.tolist() == [0, 0, 0, 0], f"results differing from the operators (div, div x3, rcp, sqrt): {mism.tolist()}"
    # the fallback must actually be exercised by the extreme classes, and must not be the common case
    assert all(f > 0 for f in fallbacks.tolist())
    assert all(f < 0.75 * (1 << 27) for f in fallbacks.tolist())


def test_fp64_probe_reports_sane_rates():
    r = admm_b200.probe_fp64()
    # 64 FP64 lanes per SM: DFMA issue rate = sms * 64 * clock
    nominal = r["sms"] * 64 * r["sm_max_mhz"] * 1e6 / 1e12
    for k in ("dfma_tinst_s", "dadd_tinst_s", "dmul_tinst_s"):
        assert 0.3 * nominal < r[k] < 1.15 * nominal, (k, r[k], nominal)
    assert 2.0 < r["dfma_dependent_cycles"] < 40.0
