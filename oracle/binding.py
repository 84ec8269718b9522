"""ctypes binding of the CPU oracle (oracle/libida_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package rust_ida_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libida_oracle.so")

PK_ROBERTS, PK_LORENZ63, PK_HEAT2D = 0, 1, 2
COUNTER_NAMES = ("nst", "nre", "nje", "nni", "netf", "ncfn", "nsetups", "nls_nconvfails", "nattempts")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
POW_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)


def build(force=False):
    src = os.path.join(_HERE, "ida_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src) or not os.path.exists(os.path.join(_HERE, "libida_oracle_ops.so")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_pow.restype = C.c_double
        L.oracle_pow.argtypes = [C.c_double, C.c_double]
        L.oracle_set_pow.argtypes = [C.c_void_p]
        L.oracle_getrf.restype = C.c_int
        L.oracle_getrf.argtypes = [_dp, _i64p, C.c_int]
        L.oracle_getrs.argtypes = [_dp, _i64p, _dp, C.c_int]
        L.oracle_wrms.restype = C.c_double
        L.oracle_wrms.argtypes = [_dp, _dp, C.c_int]
        L.oracle_wrms_masked.restype = C.c_double
        L.oracle_wrms_masked.argtypes = [_dp, _dp, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"), C.c_int]
        L.oracle_res.argtypes = [C.c_int, C.c_int, _dp, C.c_double, _dp, _dp, _dp]
        L.oracle_jac.argtypes = [C.c_int, C.c_int, _dp, C.c_double, _dp, _dp, _dp]
        L.oracle_unit_new.restype = C.c_void_p
        L.oracle_unit_new.argtypes = [C.c_int]
        L.oracle_unit_free.argtypes = [C.c_void_p]
        L.oracle_unit_set_d.argtypes = [C.c_void_p, C.c_char_p, _dp, C.c_int]
        L.oracle_unit_get_d.argtypes = [C.c_void_p, C.c_char_p, _dp, C.c_int]
        L.oracle_unit_set_i.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        L.oracle_unit_get_i.restype = C.c_long
        L.oracle_unit_get_i.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_unit_set_coeffs.restype = C.c_double
        L.oracle_unit_set_coeffs.argtypes = [C.c_void_p]
        L.oracle_unit_predict.argtypes = [C.c_void_p]
        L.oracle_unit_restore.argtypes = [C.c_void_p, C.c_double]
        L.oracle_unit_test_error.restype = C.c_int
        L.oracle_unit_test_error.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_unit_complete_step.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.oracle_unit_get_solution.restype = C.c_int
        L.oracle_unit_get_solution.argtypes = [C.c_void_p, C.c_double]
        L.oracle_unit_get_dky.restype = C.c_int
        L.oracle_unit_get_dky.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp]
        L.oracle_ensemble_dky.restype = C.c_int
        L.oracle_ensemble_dky.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_int,
                                          C.c_int, C.c_int, _dp, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"),
                                          _dp, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
        L.oracle_newton_quad3.restype = C.c_int
        L.oracle_newton_quad3.argtypes = [_dp, _dp, C.c_double, C.c_int, _dp, C.POINTER(C.c_long),
                                          C.POINTER(C.c_long), C.c_int]
        L.oracle_ensemble_solve.restype = C.c_int
        L.oracle_ensemble_solve.argtypes = [
            C.c_int, C.c_int, C.c_int64, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_single_trace.restype = C.c_int
        L.oracle_single_trace.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_int, _dp, _i64p]
        _lib = L
    return _lib


def set_pow(fn_ptr):
    """Inject a pow(x, y) implementation (a C function pointer as int/c_void_p); None -> libm pow."""
    lib().oracle_set_pow(fn_ptr)


def set_id(id_, suppressalg=True):
    """ida_id / ida_suppressalg for the ensemble entry points that follow (process-global); set_id(None) switches it off"""
    L = lib()
    L.oracle_set_id.argtypes = [C.c_void_p, C.c_int, C.c_int]
    if id_ is None:
        L.oracle_set_id(None, 0, 0)
        return
    m = np.ascontiguousarray(np.asarray(id_).reshape(-1) != 0, dtype=np.uint8)
    L.oracle_set_id(m.ctypes.data, m.size, 1 if suppressalg else 0)


def getrf(a):
    """a: (n, n) mathematical matrix.  Returns (LU as (n,n) mathematical, pivots int64, info)."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    col = np.ascontiguousarray(a.T).copy()  # column-major storage
    piv = np.zeros(n, dtype=np.int64)
    info = lib().oracle_getrf(col, piv, n)
    return col.T.copy(), piv, info


def getrs(lu, piv, b):
    lu = np.asarray(lu, dtype=np.float64)
    n = lu.shape[0]
    col = np.ascontiguousarray(lu.T).copy()
    x = np.array(b, dtype=np.float64).copy()
    lib().oracle_getrs(col, np.ascontiguousarray(piv, dtype=np.int64), x, n)
    return x


def wrms(x, w):
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    return lib().oracle_wrms(x, w, x.size)


def wrms_masked(x, w, id_):
    """src/norm_rms.rs:49-57"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    m = np.ascontiguousarray(np.asarray(id_) != 0, dtype=np.uint8)
    return lib().oracle_wrms_masked(x, w, m, x.size)


def res(kind, params, yy, yp, t=0.0):
    yy = np.ascontiguousarray(yy, dtype=np.float64)
    r = np.zeros_like(yy)
    lib().oracle_res(kind, yy.size, np.ascontiguousarray(params, dtype=np.float64), t, yy,
                     np.ascontiguousarray(yp, dtype=np.float64), r)
    return r


def jac(kind, params, cj, yy, yp):
    yy = np.ascontiguousarray(yy, dtype=np.float64)
    n = yy.size
    J = np.zeros((n, n))
    lib().oracle_jac(kind, n, np.ascontiguousarray(params, dtype=np.float64), cj, yy,
                     np.ascontiguousarray(yp, dtype=np.float64), J)
    return J.T.copy()  # column-major storage -> mathematical (i, j)


class Unit:
    """One oracle Ida whose private state can be poked, like src/tests/mod.rs does with `Dummy`."""

    def __init__(self, n=3):
        self.n = n
        self.h = lib().oracle_unit_new(n)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_unit_free(self.h)
            self.h = None

    def set(self, field, value):
        if isinstance(value, (int, np.integer)) and field in ("kk", "kused", "knew", "phase", "ns", "nst", "maxord"):
            assert lib().oracle_unit_set_i(self.h, field.encode(), int(value)) == 0, field
            return
        v = np.ascontiguousarray(np.asarray(value, dtype=np.float64).reshape(-1))
        assert lib().oracle_unit_set_d(self.h, field.encode(), v, v.size) == 0, field

    def get(self, field, shape=None):
        if field in ("kk", "kused", "knew", "phase", "ns", "nst", "maxord"):
            return int(lib().oracle_unit_get_i(self.h, field.encode()))
        size = {"phi": 6 * self.n, "psi": 6, "alpha": 6, "beta": 6, "sigma": 6, "gamma": 6}.get(field)
        if size is None:
            size = self.n if field in ("ee", "ewt", "delta", "yy", "yp", "yypredict", "yppredict") else 1
        out = np.zeros(size)
        assert lib().oracle_unit_get_d(self.h, field.encode(), out, size) == 0, field
        if field == "phi":
            return out.reshape(6, self.n)
        return out if size > 1 else float(out[0])

    def set_coeffs(self):
        return lib().oracle_unit_set_coeffs(self.h)

    def predict(self):
        lib().oracle_unit_predict(self.h)

    def restore(self, saved_t):
        lib().oracle_unit_restore(self.h, saved_t)

    def test_error(self, ck):
        a, b = C.c_double(), C.c_double()
        conv = lib().oracle_unit_test_error(self.h, ck, C.byref(a), C.byref(b))
        return a.value, b.value, bool(conv)

    def complete_step(self, err_k, err_km1):
        lib().oracle_unit_complete_step(self.h, err_k, err_km1)

    def get_solution(self, t):
        return lib().oracle_unit_get_solution(self.h, t)

    def get_dky(self, t, k):
        """Ida::get_dky (src/lib.rs:424-529): (status, dky[n])"""
        out = np.zeros(self.n)
        r = lib().oracle_unit_get_dky(self.h, t, k, out)
        return r, out


def newton_quad3(y0, w, tol, maxiters, call_lsetup=True):
    """Newton::solve on the reference's test problem; returns (status: 0 ok, 1 ConvergenceRecover, -col singular
    Jacobian column, y, niters, nconvfails)"""
    y = np.zeros(3)
    ni, nc = C.c_long(), C.c_long()
    r = lib().oracle_newton_quad3(np.ascontiguousarray(y0, dtype=np.float64), np.ascontiguousarray(w, dtype=np.float64),
                                  tol, maxiters, y, C.byref(ni), C.byref(nc), 1 if call_lsetup else 0)
    return r, y, ni.value, nc.value


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ensemble_solve(kind, n, params, yy0, yp0, rtol, atol, touts, nthreads=1, want_traj=False):
    """params (P,B), yy0/yp0 (n,B) SoA.  Returns dict with yy, yp (n,B), counters {name: (B,)},
    kused, kk, hused, hh, tn, tret, status (ntout,B), and yy_traj (ntout,n,B) if requested."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    yy0 = np.ascontiguousarray(yy0, dtype=np.float64)
    yp0 = np.ascontiguousarray(yp0, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    touts = np.ascontiguousarray(touts, dtype=np.float64)
    P, B = params.shape
    assert yy0.shape == (n, B) and yp0.shape == (n, B) and atol.shape == (n,)
    yy = np.zeros((n, B))
    yp = np.zeros((n, B))
    traj = np.zeros((len(touts), n, B)) if want_traj else None
    counters = np.zeros((9, B), dtype=np.int64)
    si = np.zeros((2, B), dtype=np.int64)
    sd = np.zeros((4, B))
    status = np.zeros((len(touts), B), dtype=np.int32)
    lib().oracle_ensemble_solve(kind, n, B, P, params, yy0, yp0, rtol, atol, touts, len(touts), int(nthreads),
                                _ptr(yy), _ptr(yp), _ptr(traj), _ptr(counters), _ptr(si), _ptr(sd), _ptr(status))
    out = {"yy": yy, "yp": yp, "status": status, "kused": si[0], "kk": si[1], "hused": sd[0], "hh": sd[1],
           "tn": sd[2], "tret": sd[3], "counters": {k: counters[i] for i, k in enumerate(COUNTER_NAMES)}}
    if want_traj:
        out["yy_traj"] = traj
    return out


def ensemble_dky(kind, n, params, yy0, yp0, rtol, atol, touts, tq, kq, nthreads=1):
    """Integrate over `touts`, then get_dky(tq[q], kq[q]) on every instance's final state: (dky (nq,n,B), status (nq,B));
    rows of failed queries stay NaN."""
    params, yy0, yp0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (params, yy0, yp0))
    B = yy0.shape[1]
    tq = np.ascontiguousarray(tq, dtype=np.float64)
    kq = np.ascontiguousarray(kq, dtype=np.int32)
    dky = np.full((len(tq), n, B), np.nan)
    st = np.zeros((len(tq), B), dtype=np.int32)
    lib().oracle_ensemble_dky(kind, n, B, params.shape[0], params, yy0, yp0, rtol, np.ascontiguousarray(atol, dtype=np.float64),
                              np.ascontiguousarray(touts, dtype=np.float64), len(touts), nthreads, len(tq), tq, kq, dky, st)
    return dky, st


def attempt_trace(kind, params, yy0, yp0, rtol, atol, touts, max_rows=4096):
    """state at the top of every step attempt (the reference's data_trace frames, src/lib.rs:635-639) in the engine's
    field order (include/b200ida.h, b200ida_set_trace).  Returns (doubles (attempts, 10n+45+n^2), ints (attempts, 16+n), status)."""
    yy0 = np.ascontiguousarray(yy0, dtype=np.float64)
    n = yy0.size
    touts = np.ascontiguousarray(touts, dtype=np.float64)
    d = np.zeros((max_rows, 10 * n + 45 + n * n))
    q = np.zeros((max_rows, 16 + n), dtype=np.int32)
    st = np.zeros(touts.size, dtype=np.int32)
    L = lib()
    L.oracle_attempt_trace.restype = C.c_long
    L.oracle_attempt_trace.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_int, _dp,
                                       np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"), C.c_long,
                                       np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
    c = L.oracle_attempt_trace(kind, n, np.ascontiguousarray(params, dtype=np.float64), yy0,
                               np.ascontiguousarray(yp0, dtype=np.float64), rtol, np.ascontiguousarray(atol, dtype=np.float64),
                               touts, touts.size, d, q, max_rows, st)
    if c > max_rows:
        raise RuntimeError("trace overflow: %d attempts" % c)
    return d[:c], q[:c], st


def roots_run(kind, params, yy0, yp0, rtol, atol, touts, max_ev=64):
    """one instance with rootfinding, driven like examples/roberts.rs (solve again after every Root return).
    Returns dict: t (nev,), iroots (nev, nr), y (nev, n), yy (n,), counters {.., nge}, status (ntout,)."""
    yy0 = np.ascontiguousarray(yy0, dtype=np.float64)
    n = yy0.size
    nr = 2 if kind == PK_ROBERTS else (1 if kind == PK_LORENZ63 else 0)
    touts = np.ascontiguousarray(touts, dtype=np.float64)
    ev_t = np.zeros(max_ev)
    ev_r = np.zeros((max_ev, max(nr, 1)))
    ev_y = np.zeros((max_ev, n))
    yy = np.zeros(n)
    cnt = np.zeros(10, dtype=np.int64)
    st = np.zeros(touts.size, dtype=np.int32)
    L = lib()
    L.oracle_roots_run.restype = C.c_long
    L.oracle_roots_run.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_int, C.c_long, _dp, _dp, _dp, _dp, _i64p,
                                   np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
    nev = L.oracle_roots_run(kind, n, np.ascontiguousarray(params, dtype=np.float64), yy0, np.ascontiguousarray(yp0, dtype=np.float64),
                             rtol, np.ascontiguousarray(atol, dtype=np.float64), touts, touts.size, max_ev, ev_t, ev_r, ev_y, yy, cnt, st)
    if nev > max_ev:
        raise RuntimeError("more than %d root events" % max_ev)
    names = COUNTER_NAMES + ("nge",)
    return {"t": ev_t[:nev], "iroots": ev_r[:nev, :nr], "y": ev_y[:nev], "yy": yy, "status": st,
            "counters": {k: int(cnt[i]) for i, k in enumerate(names)}}


def single_trace(kind, params, yy0, yp0, rtol, atol, touts):
    yy0 = np.ascontiguousarray(yy0, dtype=np.float64)
    n = yy0.size
    touts = np.ascontiguousarray(touts, dtype=np.float64)
    rows = np.zeros((len(touts), 3 + n))
    counters = np.zeros(9, dtype=np.int64)
    st = lib().oracle_single_trace(kind, n, np.ascontiguousarray(params, dtype=np.float64), yy0,
                                   np.ascontiguousarray(yp0, dtype=np.float64), rtol,
                                   np.ascontiguousarray(atol, dtype=np.float64), touts, len(touts), rows, counters)
    return st, rows, dict(zip(COUNTER_NAMES, counters.tolist()))


_ops_lib = None


def ensemble_ops(kind, n, params, yy0, yp0, rtol, atol, touts, nthreads=1):
    """Algorithmic FP64 operation counts of the reference algorithm (oracle built with -DORACLE_COUNT_OPS,
    libida_oracle_ops.so): returns (ops (B, 5) int64 = add, mul, div, sqrt, pow per instance, attempts (B,) int64)."""
    global _ops_lib
    if _ops_lib is None:
        build()
        _ops_lib = C.CDLL(os.path.join(_HERE, "libida_oracle_ops.so"))
        _ops_lib.oracle_ensemble_ops.restype = C.c_int
        _ops_lib.oracle_ensemble_ops.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp, C.c_int,
                                                 C.c_int, _i64p, _i64p]
    params, yy0, yp0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (params, yy0, yp0))
    B = yy0.shape[1]
    ops = np.zeros((B, 5), dtype=np.int64)
    natt = np.zeros(B, dtype=np.int64)
    touts = np.ascontiguousarray(touts, dtype=np.float64)
    _ops_lib.oracle_ensemble_ops(kind, n, B, params.shape[0], params, yy0, yp0, rtol, np.ascontiguousarray(atol, dtype=np.float64),
                                 touts, len(touts), int(nthreads), ops.reshape(-1), natt)
    return ops, natt
