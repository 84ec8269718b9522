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
