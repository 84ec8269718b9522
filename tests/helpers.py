"""Shared helpers for the parity tests."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROBERTS_ATOL = np.array([1e-8, 1e-6, 1e-6])
ROBERTS_RTOL = 1e-4
ROBERTS_TOUTS = 0.4 * 10.0 ** np.arange(12)
LORENZ_TOUTS = 0.5 * np.arange(1, 11)
LORENZ_RTOL, LORENZ_ATOL = 1e-6, np.full(3, 1e-8)


def detpow_host_lib():
    """host build of the engine's pow (csrc/detpow.h), used to give the oracle the same pow"""
    path = os.path.join(ROOT, "rust_ida_b200", "csrc", "libdetpow_host.so")
    L = C.CDLL(path)
    L.detpow_host.restype = C.c_double
    L.detpow_host.argtypes = [C.c_double, C.c_double]
    L.detpow_compare_libm.restype = C.c_int64
    L.detpow_compare_libm.argtypes = [C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p]
    L.detpow_host_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    return L


def parity_oracle():
    """oracle.binding with its DEFAULT pow: the platform libm `pow`, i.e. what the reference's f64::powf calls.  Every
    GPU parity test compares the device results with this oracle, so a det_pow / libm mismatch on a real trajectory
    fails a test instead of hiding behind a shared pow."""
    from oracle import binding as ob
    ob.set_pow(None)
    return ob, None


def oracle_with_detpow():
    """oracle.binding with pow switched to the engine's det_pow (csrc/detpow.h built for the host).  Only used by
    tests/test_detpow.py to show that the two oracles agree on whole trajectories; tests/conftest.py restores the libm
    pow after every test, so the swap never leaks into another test."""
    from oracle import binding as ob
    L = detpow_host_lib()
    ob.set_pow(C.cast(L.detpow_host, C.c_void_p))
    return ob, L


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ba, bb = bits(a), bits(b)
    bad = np.nonzero(ba != bb)
    assert bad[0].size == 0, "%s: %d of %d values differ; first at %s: %r vs %r" % (
        what, bad[0].size, ba.size, tuple(x[0] for x in bad), a[tuple(x[0] for x in bad)], b[tuple(x[0] for x in bad)])
