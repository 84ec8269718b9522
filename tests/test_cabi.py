"""C-ABI checks that need no GPU: the library loads, exports every symbol include/b200ida.h declares,
and compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rust_ida_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200ida.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ida_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), "libb200ida.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names  # the Python binding covers exactly the header


def test_version():
    assert _lib.lib().b200ida_version() == 100


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "b200ida.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", hdr, flags=re.S).lower()
    assert "at::" not in hdr and "std::" not in hdr


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU error path")
def test_compute_fails_loudly_without_gpu():
    L = _lib.lib()
    a = np.eye(3).reshape(3, 3, 1).copy()
    piv = np.zeros((3, 1), dtype=np.int64)
    info = np.zeros(1, dtype=np.int32)
    rc = L.b200ida_getrf_batched(a.ctypes.data, piv.ctypes.data, info.ctypes.data, 3, 1, 0, _lib.MEM_HOST, None)
    assert rc == _lib.ERR_NO_DEVICE
    assert b"no CUDA device" in L.b200ida_last_error()
    h = C.c_void_p()
    p = np.array([[0.04], [1e4], [3e7]])
    y = np.array([[1.0], [0.0], [0.0]])
    atol = np.array([1e-8, 1e-6, 1e-6])
    rc = L.b200ida_create(C.byref(h), 0, 3, 1, 3, p.ctypes.data, y.ctypes.data, y.ctypes.data, 1e-4, atol.ctypes.data,
                          _lib.MEM_HOST, 0, None)
    assert rc == _lib.ERR_NO_DEVICE
    import rust_ida_b200 as r
    with pytest.raises(r.B200IdaError):
        r.Ida(r.Roberts(p), y, y.copy(), r.TolControlSV(1e-4, atol))
    with pytest.raises(r.B200IdaError):
        r.Dense.new(3, 1).setup(a)


def test_argument_errors():
    L = _lib.lib()
    rc = L.b200ida_getrf_batched(None, None, None, 3, 1, 0, _lib.MEM_HOST, None)
    assert rc == _lib.ERR_ARG
    a = np.zeros(4)
    rc = L.b200ida_wrms_batched(a.ctypes.data, a.ctypes.data, a.ctypes.data, 300, 1, 0, _lib.MEM_HOST, None)
    assert rc == _lib.ERR_ARG
