"""`Ida` -- the reference's stepping API (Ida::new / Ida::solve / getters) over an ensemble.

Mirrors src/lib.rs:278-283 (new), src/impl_solve.rs:69-74 (solve) and src/ida_io.rs (getters): same
names, argument meaning and statuses, but every call acts on `batch` independent instances at once and
all per-step work happens inside one CUDA kernel launch (C ABI "Level 2", include/b200ida.h).
"""
import ctypes as C
import enum

import numpy as np

from . import _buffers as B
from . import _lib
from .tol_control import TolControlSV


class IdaTask(enum.Enum):  # src/lib.rs:51-55
    Normal = 0
    OneStep = 1


class IdaSolveStatus(enum.IntEnum):
    """0 = IdaSolveStatus::Success (src/lib.rs:57-63); negatives = the IdaError the reference returns."""
    Success = 0
    TooMuchWork = -1      # mxstep steps in one call (IdaError::IllegalInput, src/impl_solve.rs:249-260)
    TooMuchAccuracy = -2  # src/impl_solve.rs:296-308
    ErrFail = -3          # src/lib.rs:1194-1198
    ConvergenceFail = -4  # src/lib.rs:1232-1234
    LinearSetupFail = -5  # singular Jacobian (reference: panic, src/ida_ls.rs:287)
    BadErrorWeight = -6   # src/impl_solve.rs:272-285
    IllegalInput = -7     # src/impl_solve.rs:97-109
    BadTimeValue = -8     # src/lib.rs:1286-1292
    BadK = -10            # IdaError::BadK from get_dky, src/lib.rs:443-447
    CloseRoots = -9       # IdaError::CloseRoots, src/impl_r_check.rs:199-203
    Root = 2              # IdaSolveStatus::Root (src/lib.rs:57-63); consumed inside a ladder launch (enable_roots)


def set_default_reorder(mode):
    """"off" | "on" | "auto": internal warp-coherence ordering of instances (include/b200ida.h); results and
    the caller-visible instance order are unaffected."""
    _lib.check(_lib.lib().b200ida_set_default_reorder({"off": 0, "on": 1, "auto": 2}[mode]))


class AttemptTrace:
    """device buffers of Ida.trace(); rows() -> (doubles (attempts, 10n+45+n^2), ints (attempts, 16+n)) as numpy"""

    def __init__(self, ida, d, i, count):
        self.ida, self.d, self.i, self.count = ida, d, i, count

    def rows(self):
        self.ida.sync()
        c = int(self.count.cpu()[0])
        if c > self.d.shape[0]:
            raise RuntimeError("trace overflow: %d attempts, room for %d" % (c, self.d.shape[0]))
        return self.d[:c].cpu().numpy(), self.i[:c].cpu().numpy()


class RootEvents:
    """device buffers of Ida.enable_roots(): t (max_events, batch), iroots (max_events, 2, batch) = get_root_info at
    each event (src/ida_io.rs:115), count (batch,), nge (batch,) = get_num_g_evals (src/ida_io.rs:107)"""

    def __init__(self, ida, t, iroots, count, nge):
        self.ida, self.t, self.iroots, self.count, self.nge = ida, t, iroots, count, nge

    def numpy(self):
        self.ida.sync()
        return {"t": self.t.cpu().numpy(), "iroots": self.iroots.cpu().numpy(), "count": self.count.cpu().numpy(),
                "nge": self.nge.cpu().numpy()}


class Ida:
    """Ida<P, Dense, Newton, TolControlSV> for a whole ensemble."""

    def __init__(self, problem, yy0, yp0, tol_control, device=None):
        """Ida::new(problem, yy0, yp0, tol_control) (src/lib.rs:278-283).  yy0, yp0: (n, batch) SoA,
        numpy (host; copied in) or torch CUDA tensors (device), same space as problem.params.
        device: CUDA device of the ensemble.  For CUDA tensors it is taken from the tensors (which must all live on one
        device; an explicit `device` that disagrees is an error); for host buffers it defaults to 0."""
        _lib.require_device()
        tdev = B.device_of(problem.params, yy0, yp0)
        if tdev is not None:
            if device is not None and int(device) != tdev:
                raise ValueError("device=%r but params / yy0 / yp0 live on cuda:%d" % (device, tdev))
            device = tdev
        device = 0 if device is None else int(device)
        n, batch = problem.model_size(), problem.batch
        if tuple(yy0.shape) != (n, batch) or tuple(yp0.shape) != (n, batch):
            raise ValueError("yy0 / yp0 must have shape (%d, %d)" % (n, batch))  # assert_eq!, src/lib.rs:284
        if not isinstance(tol_control, TolControlSV):
            raise TypeError("tol_control must be TolControlSV or TolControlSS")
        self.problem, self.n, self.batch, self.device = problem, n, batch, device
        self._atol = np.ascontiguousarray(tol_control.atol_vector(n))
        pp, m1 = B.describe(problem.params, np.float64, "params")
        py, m2 = B.describe(yy0, np.float64, "yy0")
        pyp, m3 = B.describe(yp0, np.float64, "yp0")
        self._mem = B.common_mem(m1, m2, m3)
        self._ref = yy0  # which space outputs are allocated in
        h = C.c_void_p()
        with B.on_device(device):
            _lib.check(_lib.lib().b200ida_create(C.byref(h), problem.kind, n, batch, problem.nparams, pp, py, pyp,
                                                 tol_control.ida_rtol, self._atol.ctypes.data, self._mem, device,
                                                 B.current_stream(self._mem, device)))
        self._h = h

    new = classmethod(lambda cls, *a, **k: cls(*a, **k))

    def _call(self, name, *args):
        """one C-ABI call on this ensemble's device; the caller's current CUDA device is restored afterwards"""
        with B.on_device(self.device):
            _lib.check(getattr(_lib.lib(), name)(*args))

    def close(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib().b200ida_destroy(self._h)
            except Exception:  # interpreter shutdown: the module globals may be gone already
                pass
            self._h = None

    __del__ = close

    KERNELS = {"auto": 0, "thread": 1, "warp": 2, "cta": 3}

    def set_kernel(self, kernel):
        """choose the kernel family ("auto", "thread", "warp", "cta"); results are bit-identical"""
        self._call("b200ida_set_kernel", self._h, self.KERNELS[kernel])

    def set_regroup(self, trips_per_segment=96, max_segments=4):
        """engine tuning (results are bit-identical): segment length and count of the re-grouped production solve;
        max_segments <= 1 switches re-grouping off"""
        self._call("b200ida_set_regroup", self._h, int(trips_per_segment), int(max_segments))

    def set_id(self, id_, suppressalg=True):
        """ida_id / ida_suppressalg (src/lib.rs:102,122): id_[i] truthy = differential component, falsy = algebraic.
        With suppressalg the local-error norms leave the algebraic components out (norm_wrms_masked,
        src/norm_rms.rs:49-57).  set_id(None) restores the reference's defaults."""
        if id_ is None:
            self._call("b200ida_set_id", self._h, None, 0)
            return
        m = np.ascontiguousarray(np.asarray(id_).reshape(-1) != 0, dtype=np.uint8)
        if m.size != self.n:
            raise ValueError("id has %d entries, the systems have %d components" % (m.size, self.n))
        self._call("b200ida_set_id", self._h, m.ctypes.data, 1 if suppressalg else 0)

    def reinit(self, yy0, yp0, params=None):
        """fresh Ida::new state for every instance (t = 0, counters 0), keeping the device allocation"""
        py, m1 = B.describe(yy0, np.float64, "yy0")
        pyp, m2 = B.describe(yp0, np.float64, "yp0")
        pp, m3 = B.describe(params, np.float64, "params")
        mem = B.common_mem(m1, m2, m3)
        self._call("b200ida_reinit", self._h, pp, py, pyp, mem)

    # ---- stepping ---------------------------------------------------------------------------------------
    def solve(self, tout, itask=IdaTask.Normal):
        """Ida::solve(tout, &mut tret, itask) (src/impl_solve.rs:69-74) for every instance.
        Returns (status, tret): int32 (batch,) of IdaSolveStatus values and float64 (batch,)."""
        if itask != IdaTask.Normal:
            raise NotImplementedError("only IdaTask::Normal is on the device path")
        status = B.empty_like_space(self._ref, (self.batch,), np.int32)
        ps, _ = B.describe(status, np.int32, "status")
        self._call("b200ida_solve", self._h, float(tout), 0, ps, self._mem)
        tret = B.empty_like_space(self._ref, (self.batch,), np.float64)
        pt, _ = B.describe(tret, np.float64, "tret")
        self._call("b200ida_get_state", self._h, None, None, pt, self._mem)
        return status, tret

    def solve_ladder(self, touts, itask=IdaTask.Normal, want_traj=True):
        """solve() for touts[0], touts[1], ... in ONE launch.  Returns (status (ntout, batch),
        yy_traj (ntout, n, batch) or None)."""
        if itask != IdaTask.Normal:
            raise NotImplementedError("only IdaTask::Normal is on the device path")
        touts = np.ascontiguousarray(touts, dtype=np.float64)
        nt = touts.size
        status = B.empty_like_space(self._ref, (nt, self.batch), np.int32)
        traj = B.empty_like_space(self._ref, (nt, self.n, self.batch), np.float64) if want_traj else None
        ps, _ = B.describe(status, np.int32, "status")
        pj, _ = B.describe(traj, np.float64, "yy_traj")
        self._call("b200ida_solve_ladder", self._h, touts.ctypes.data, nt, 0, pj, ps, self._mem)
        return status, traj

    def fallback_count(self):
        """instances the last solve handed from the fast-path-only pass to the exact pass (diagnostic; results do not
        depend on it)"""
        c = C.c_int64(0)
        self._call("b200ida_get_fallback_count", self._h, C.byref(c), None)
        return int(c.value)

    def fallback_histogram(self):
        """histogram (16 bins) of the per-instance fallback flags (reason codes in -DIDA_POISON_REASON=1 builds)"""
        c = C.c_int64(0)
        h = (C.c_int64 * 16)()
        self._call("b200ida_get_fallback_count", self._h, C.byref(c), h)
        return [int(v) for v in h]

    def sync(self):
        self._call("b200ida_sync", self._h)

    # ---- rootfinding: src/impl_r_check.rs, Root returns of Ida::solve, get_root_info / get_num_g_evals ----------------
    def enable_roots(self, max_events=16):
        """Switch rootfinding on (Roberts: g0 = y1 - 1e-4, g1 = y3 - 0.01).  Every following solve_ladder launch logs
        the Root returns per instance like the loop of examples/roberts.rs:117-140.  Returns a RootEvents object;
        enable_roots(0) switches it off."""
        import torch
        if not max_events:
            self._call("b200ida_set_roots", self._h, 0, None, None, None, None)
            self._roots = None
            return None
        dev = torch.device("cuda", self.device)
        nr = int(getattr(self.problem, "nroots", 0))
        if nr < 1:
            raise NotImplementedError("%s has no root functions (Root::num_roots == 0)" % type(self.problem).__name__)
        ev = RootEvents(self, torch.zeros((max_events, self.batch), dtype=torch.float64, device=dev),
                        torch.zeros((max_events, nr, self.batch), dtype=torch.float64, device=dev),
                        torch.zeros((self.batch,), dtype=torch.int32, device=dev),
                        torch.zeros((self.batch,), dtype=torch.int32, device=dev))
        self._call("b200ida_set_roots", self._h, int(max_events), ev.t.data_ptr(), ev.iroots.data_ptr(),
                                                ev.count.data_ptr(), ev.nge.data_ptr())
        self._roots = ev  # keeps the device buffers alive while the engine holds their addresses
        return ev

    # ---- per-attempt state trace (the reference's `data_trace` feature, src/lib.rs:297-301, 635-639) --------------
    def trace(self, instance, max_attempts=4096):
        """Record the complete solver state of `instance` at the top of every step attempt of the following solves
        (thread-per-instance kernels).  Returns an AttemptTrace; call .rows() after the solve; trace(None) stops."""
        import torch
        if instance is None:
            self._call("b200ida_set_trace", self._h, -1, None, None, None, 0)
            self._trace = None
            return None
        dev = torch.device("cuda", self.device)
        nd, ni = 10 * self.n + 45 + self.n * self.n, 16 + self.n
        t = AttemptTrace(self, torch.zeros((max_attempts, nd), dtype=torch.float64, device=dev),
                         torch.zeros((max_attempts, ni), dtype=torch.int32, device=dev),
                         torch.zeros((1,), dtype=torch.int32, device=dev))
        self._call("b200ida_set_trace", self._h, int(instance), t.d.data_ptr(), t.i.data_ptr(), t.count.data_ptr(),
                                                int(max_attempts))
        self._trace = t  # keeps the device buffers alive while the engine holds their addresses
        return t

    # ---- getters: src/ida_io.rs -----------------------------------------------------------------------------
    def _state(self, which):
        out = B.empty_like_space(self._ref, (self.n, self.batch) if which < 2 else (self.batch,), np.float64)
        p, _ = B.describe(out, np.float64, "out")
        args = [None, None, None]
        args[which] = p
        self._call("b200ida_get_state", self._h, args[0], args[1], args[2], self._mem)
        return out

    def get_yy(self):  # ida_io.rs:11
        return self._state(0)

    def get_yp(self):  # ida_io.rs:16
        return self._state(1)

    def _counter(self, idx):
        out = B.empty_like_space(self._ref, (self.batch,), np.int64)
        p, _ = B.describe(out, np.int64, "out")
        args = [None] * 8
        args[idx] = p
        self._call("b200ida_get_counters", self._h, *args, self._mem)
        return out

    def get_num_steps(self):  # :58
        return self._counter(0)

    def get_num_res_evals(self):  # :64
        return self._counter(1)

    def get_num_jac_evals(self):  # :79
        return self._counter(2)

    def get_num_nonlin_solv_iters(self):  # :84
        return self._counter(3)

    def get_num_err_test_fails(self):  # :74
        return self._counter(4)

    def get_num_nonlin_solv_conv_fails(self):  # :104
        return self._counter(5)

    def get_num_lin_solv_setups(self):  # :69
        return self._counter(6)

    def get_num_step_attempts(self):  # data_trace frames, src/lib.rs:635-639
        return self._counter(7)

    def get_num_lin_res_evals(self):  # :92 (always 0: no DQ Jacobian)
        return B.empty_like_space(self._ref, (self.batch,), np.int64) * 0

    def get_dky(self, t, k, out=None):
        """Ida::get_dky (src/lib.rs:424-529) on every instance: the k-th derivative at time `t` of each instance's
        interpolating polynomial over its last step.  Returns (dky (n, batch), status (batch,) int32: 0 / BadK /
        BadTimeValue); columns of failed instances keep what `out` held (NaN when `out` is not given)."""
        if out is None:
            out = B.empty_like_space(self._ref, (self.n, self.batch), np.float64)
            out[...] = float("nan")
        st = B.empty_like_space(self._ref, (self.batch,), np.int32)
        po, _ = B.describe(out, np.float64, "out")
        ps, _ = B.describe(st, np.int32, "status")
        self._call("b200ida_get_dky", self._h, float(t), int(k), po, ps, self._mem)
        return out, st

    def _stepinfo(self, idx):
        dt = np.int32 if idx < 2 else np.float64
        out = B.empty_like_space(self._ref, (self.batch,), dt)
        p, _ = B.describe(out, dt, "out")
        args = [None] * 6
        args[idx] = p
        self._call("b200ida_get_step_info", self._h, *args, self._mem)
        return out

    def get_last_order(self):  # :22
        return self._stepinfo(0)

    def get_current_order(self):  # :27
        return self._stepinfo(1)

    def get_last_step(self):  # :38
        return self._stepinfo(2)

    def get_current_setp(self):  # :43 (sic, the reference's spelling)
        return self._stepinfo(3)

    get_current_step = get_current_setp

    def get_current_time(self):  # :48
        return self._stepinfo(4)

    def get_actual_init_step(self):  # :32
        return self._stepinfo(5)

    def counters(self):
        names = ("nst", "nre", "nje", "nni", "netf", "ncfn", "nsetups", "nattempts")
        return {k: self._counter(i) for i, k in enumerate(names)}
