"""`Newton` -- the reference's nonlinear solver (crates/nonlinear) over a batch of independent problems.

Mirrors `NLSolver<T, D>` (crates/nonlinear/src/traits.rs:129-209) as implemented by `Newton<T, D>`
(crates/nonlinear/src/newton.rs:14-167): `Newton.new(maxiters)`, `solve(problem, y0, y, w, tol, call_lsetup)`,
`get_num_iters()`, `get_cur_iter()`, `get_num_conv_fails()` -- same names and argument meaning, but every call acts on
`batch` systems at once (C ABI "Level 1.5", b200ida_newton_batched).  The reference's `Result` becomes a per-system
status array; `NLProblem`s are device functors selected by a kind (`Quad3Problem` = the reference's own test problem).
"""
import numpy as np

from . import _buffers as B
from . import _lib

NLP_QUAD3 = 0


class ConvergenceRecover(RuntimeError):
    """nonlinear::Error::ConvergenceRecover (crates/nonlinear/src/lib.rs:10-35) for at least one system"""

    def __init__(self, status):
        super().__init__("%d of %d systems did not converge" % (int((status != 0).sum()), status.size))
        self.status = status


class Quad3Problem:
    """the NLProblem of the reference's Newton test (newton.rs:192-343): x^2+y^2+z^2 = 1, 2x^2+y^2 = 4z, 3x^2+z^2 = 4y"""
    kind, n = NLP_QUAD3, 3


class Newton:
    def __init__(self, maxiters):  # NLSolver::new, traits.rs:136
        self.maxiters = int(maxiters)
        self.status = self.niters = self.nconvfails = None

    new = classmethod(lambda cls, maxiters: cls(maxiters))

    def solve(self, problem, y0, y, w, tol, call_lsetup, raise_on_failure=False):
        """NLSolver::solve (traits.rs:176-193; Newton: newton.rs:51-167) for every system: y0, y, w are (n, batch) SoA,
        numpy (host) or torch CUDA tensors (device).  Returns the status array (0 Ok, 1 ConvergenceRecover,
        -col LinearSetupFailed); the per-system counters accumulate like the reference's `niters` / `nconvfails`."""
        _lib.require_device()
        n, batch = y0.shape
        if n != problem.n or tuple(y.shape) != (n, batch) or tuple(w.shape) != (n, batch):
            raise ValueError("y0, y, w must have shape (%d, batch)" % problem.n)
        p0, m0 = B.describe(y0, np.float64, "y0")
        py, m1 = B.describe(y, np.float64, "y")
        pw, m2 = B.describe(w, np.float64, "w")
        mem = B.common_mem(m0, m1, m2)
        st = B.empty_like_space(y0, (batch,), np.int32)
        ni = B.empty_like_space(y0, (batch,), np.int32)
        nc = B.empty_like_space(y0, (batch,), np.int32)
        ctx, dev = B.tensors_device(y0, y, w)
        with ctx:
            _lib.check(_lib.lib().b200ida_newton_batched(problem.kind, n, batch, p0, pw, float(tol), self.maxiters,
                                                         1 if call_lsetup else 0, py, B.describe(st, np.int32, "st")[0],
                                                         B.describe(ni, np.int32, "ni")[0], B.describe(nc, np.int32, "nc")[0],
                                                         mem, B.current_stream(mem, dev)))
        self.status = st
        self.niters = ni if self.niters is None else self.niters + ni
        self.nconvfails = nc if self.nconvfails is None else self.nconvfails + nc
        self._last_iters = ni
        if raise_on_failure:
            s = st.cpu().numpy() if B.is_torch(st) else st
            if (s != 0).any():
                raise ConvergenceRecover(s)
        return st

    def get_num_iters(self):  # traits.rs:196
        return self.niters

    def get_cur_iter(self):  # traits.rs:202 (iterations of the last solve)
        return self._last_iters

    def get_num_conv_fails(self):  # traits.rs:205
        return self.nconvfails
