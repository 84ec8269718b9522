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


def test_pow_with_exponent_one_is_the_identity():
    """ctest's rate = (delnrm / oldnrm)^(1/m) with m == 1 (second Newton iteration, src/ida_nls.rs:243-253) is
    evaluated as the base itself on the device: libm pow(x, 1.0) -- the reference's f64::powf -- returns x
    bit-for-bit.  Checked against the platform libm on 4M arguments over the whole range ctest can produce
    (ratios of WRMS norms: non-negative, incl. 0, denormals, inf and nan) and on the engine's own det_pow."""
    L = detpow_host_lib()
    libm = C.CDLL("libm.so.6")
    libm.pow.restype = C.c_double
    libm.pow.argtypes = [C.c_double, C.c_double]
    rng = np.random.default_rng(7)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 2_000_000), rng.uniform(0.0, 2.0, 2_000_000),
                        np.array([0.0, 5e-324, 2.2e-308, 1.0, np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0), 1.7e308, np.inf])])
    out = np.empty_like(x)
    y = np.ones_like(x)
    L.detpow_host_batch(x.ctypes.data, y.ctypes.data, out.ctypes.data, len(x))
    assert np.array_equal(out.view(np.uint64), x.view(np.uint64)), "det_pow(x, 1.0) != x"
    sample = np.concatenate([x[:200_000], x[-8:]])
    ref = np.array([libm.pow(float(v), 1.0) for v in sample])
    assert np.array_equal(ref.view(np.uint64), sample.view(np.uint64)), "libm pow(x, 1.0) != x"
    assert np.isnan(libm.pow(float("nan"), 1.0))


def test_detpow_oracle_and_libm_oracle_agree_on_whole_trajectories():
    """The GPU parity tests compare with the libm-pow oracle (helpers.parity_oracle).  This closes the loop on the CPU:
    the same oracle with the engine's det_pow swapped in gives bit-identical states and counters on perturbed Roberts
    and Lorenz ensembles, i.e. det_pow == libm pow on every argument real trajectories produce."""
    from helpers import (LORENZ_ATOL, LORENZ_RTOL, LORENZ_TOUTS, ROBERTS_ATOL, ROBERTS_RTOL, ROBERTS_TOUTS, oracle_with_detpow,
                         parity_oracle)
    from rust_ida_b200.problems import Lorenz63, Roberts, lorenz_ensemble_params, roberts_ensemble_params
    for kind, P, params, rtol, atol, touts in (
            (0, Roberts, roberts_ensemble_params(3000, seed=99), ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS),
            (1, Lorenz63, lorenz_ensemble_params(1500, seed=99), LORENZ_RTOL, LORENZ_ATOL, LORENZ_TOUTS)):
        yy0, yp0 = P.consistent_ic(params)
        ob, _ = parity_oracle()
        a = ob.ensemble_solve(kind, 3, params, yy0, yp0, rtol, atol, touts, nthreads=8)
        ob, _ = oracle_with_detpow()
        b = ob.ensemble_solve(kind, 3, params, yy0, yp0, rtol, atol, touts, nthreads=8)
        assert np.array_equal(a["yy"].view(np.uint64), b["yy"].view(np.uint64))
        for k in ("nst", "nni", "nje", "netf", "ncfn"):
            assert np.array_equal(a["counters"][k], b["counters"][k]), k
