"""det_pow (csrc/detpow.h, host build) must equal the platform libm pow -- the reference's f64::powf --
bit-for-bit on the engine's argument ranges (hard part 1)."""
import ctypes as C

import numpy as np
import pytest

from helpers import detpow_host_lib

# every exponent the engine can pass: -1/(k+1), k = 1..5 (complete_step, handle_n_flag); 1/m, m = 1..3 (ctest)
EXPONENTS = np.array([-1 / 2, -1 / 3, -1 / 4, -1 / 5, -1 / 6, 1.0, 1 / 2, 1 / 3], dtype=np.float64)


@pytest.mark.parametrize("lo,hi,n", [(1e-4, 60.0, 4_000_000), (1e-8, 1e8, 2_000_000), (0.99, 1.01, 1_000_000),
                                     (1e-12, 0.9, 1_000_000)])
def test_detpow_equals_libm_pow(lo, hi, n):
    L = detpow_host_lib()
    bad = np.zeros(20)
    fb = C.c_int64()
    mism = L.detpow_compare_libm(n, 1234, lo, hi, EXPONENTS.ctypes.data, len(EXPONENTS), bad.ctypes.data, 10, C.byref(fb))
    assert fb.value == 0, "arguments left the main path"
    assert mism == 0, "det_pow != libm pow at (x, y) = %r" % (bad[:2 * min(mism, 10)].reshape(-1, 2),)


def test_detpow_known_values():
    L = detpow_host_lib()
    # rr of src/tests/complete_step.rs test3: (2*err_km1 + 1e-4)^(-1/3)
    base = 2.0 * 0.0434084688121079 + 0.0001
    assert L.detpow_host(base, -(1.0 / 3.0)) == 2.2575213239991561
    assert L.detpow_host(1.0, 0.5) == 1.0
    assert L.detpow_host(4.0, 0.5) == 2.0
    assert L.detpow_host(2.0, 1.0) == 2.0


def test_detpow_out_of_domain_falls_back():
    L = detpow_host_lib()
    assert np.isnan(L.detpow_host(-2.0, 0.5))
    assert L.detpow_host(0.0, 0.5) == 0.0
    assert L.detpow_host(np.inf, -0.5) == 0.0
