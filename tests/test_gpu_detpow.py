"""device det_pow == host det_pow (== glibc pow, tests/test_detpow.py) bit-for-bit."""
import numpy as np
import pytest

from helpers import assert_bit_equal, detpow_host_lib

pytestmark = pytest.mark.gpu


def test_device_pow_matches_host_pow():
    from rust_ida_b200 import _lib
    rng = np.random.default_rng(5)
    n = 2_000_000
    x = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), n))
    ys = np.array([-1 / 2, -1 / 3, -1 / 4, -1 / 5, -1 / 6, 1.0, 1 / 2, 1 / 3])
    y = ys[rng.integers(0, len(ys), n)]
    x[:4] = [1.0, 0.25, 1e-4, 2.0 * 0.0434084688121079 + 0.0001]
    y[:4] = [0.5, 0.5, -0.5, -1.0 / 3.0]
    out = np.empty(n)
    _lib.check(_lib.lib().b200ida_detpow_batched(x.ctypes.data, y.ctypes.data, out.ctypes.data, n, _lib.MEM_HOST, None))
    ref = np.empty(n)
    detpow_host_lib().detpow_host_batch(x.ctypes.data, y.ctypes.data, ref.ctypes.data, n)
    assert_bit_equal(out, ref, "det_pow device vs host")
    import math
    k = 200_000
    assert_bit_equal(out[:k], np.array([math.pow(a, b) for a, b in zip(x[:k], y[:k])]), "det_pow device vs libm pow")
    assert out[3] == 2.2575213239991561


def test_fastdiv_matches_division():
    """csrc/fastdiv.cuh: the straight-line division / reciprocal fast paths (batched three at a time under one range
    test, cold fallback to `/`) are bit-identical to IEEE division -- numpy's, which is the correctly rounded quotient
    the reference's `/` and `recip()` produce -- on solver-like operands, on operands across the whole exponent range
    (where the range test must send quotients to the fallback: denormal results, overflow), and on zeros / inf / nan."""
    from rust_ida_b200 import _lib
    rng = np.random.default_rng(11)
    n = 3_000_000
    a = np.concatenate([rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n),
                        rng.standard_normal(n) * 10.0 ** rng.uniform(-300, 300, n)])
    b = np.concatenate([rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n),
                        rng.standard_normal(n) * 10.0 ** rng.uniform(-300, 300, n)])
    # significands of all ones / powers of two / neighbours of 1, and the special values, in every batch position
    edge = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                     np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0), np.nextafter(2.0, 1.0), 3.0, 1.0 / 3.0, 1e-310, 4e10, 1e-8])
    ea, eb = np.meshgrid(edge, edge)
    a = np.concatenate([a, ea.ravel(), np.roll(ea.ravel(), 1), np.roll(ea.ravel(), 2)])
    b = np.concatenate([b, eb.ravel(), eb.ravel(), eb.ravel()])
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    q = np.empty_like(a)
    r = np.empty_like(a)
    _lib.check(_lib.lib().b200ida_fastdiv_batched(a.ctypes.data, b.ctypes.data, q.ctypes.data, r.ctypes.data, a.size))
    with np.errstate(all="ignore"):
        qr = a / b
        rr = 1.0 / b
    for got, ref, what in ((q, qr, "a / b"), (r, rr, "1 / b")):
        nan = np.isnan(ref)
        assert np.array_equal(np.isnan(got), nan), what
        bad = np.nonzero((got.view(np.uint64) != ref.view(np.uint64)) & ~nan)[0]
        assert bad.size == 0, "%s: %d mismatches, first a=%r b=%r got %r want %r" % (
            what, bad.size, a[bad[0]], b[bad[0]], got[bad[0]], ref[bad[0]])
