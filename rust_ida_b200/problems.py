"""Problem descriptors = the host side of the device functors in csrc/problems.cuh.

In the reference a problem implements ModelSpec / Residual / Jacobian (src/traits.rs:6-69) in Rust and
is monomorphised into `Ida<P, ..>`.  A GPU step loop cannot call back to the host per Newton iteration,
so here a problem is a `kind` (selects the device functor) plus per-instance parameter rows."""
import numpy as np

PROBLEM_ROBERTS, PROBLEM_LORENZ63, PROBLEM_HEAT2D = 0, 1, 2


class EnsembleProblem:
    kind = None
    n = None
    nparams = None

    def __init__(self, params):
        """params: (nparams, batch) float64, numpy (host) or torch CUDA tensor (device)"""
        if params.shape[0] != self.nparams:
            raise ValueError("%s needs %d parameter rows" % (type(self).__name__, self.nparams))
        self.params = params
        self.batch = int(params.shape[1])

    def model_size(self):  # ModelSpec::model_size, src/traits.rs:9
        return self.n


class UserProblem(EnsembleProblem):
    """A user-supplied Residual / Jacobian (/ Root) functor (src/traits.rs:28-89) that was compiled into the library with
    `make -C rust_ida_b200/csrc user PROBLEM_HDR=...` ("Adding a problem"); load that library through the
    B200IDA_LIB environment variable.  n, nparams and nroots must match the functor's N, NP, NR (the engine checks)."""
    kind = 3  # B200IDA_PROBLEM_USER

    def __init__(self, params, n, nroots=0):
        self.n, self.nparams, self.nroots = int(n), int(params.shape[0]), int(nroots)
        super().__init__(params)


class Roberts(EnsembleProblem):
    """Robertson kinetics (src/sample_problems/roberts.rs); params rows (k1, k2, k3), reference values
    (0.04, 1e4, 3e7)."""
    kind, n, nparams = PROBLEM_ROBERTS, 3, 3
    nroots = 2  # Root::num_roots, src/sample_problems/roberts.rs:99-113
    DEFAULT = (0.04, 1.0e4, 3.0e7)

    @classmethod
    def default(cls, batch=1):
        return cls(np.repeat(np.array(cls.DEFAULT)[:, None], batch, axis=1).copy())

    @staticmethod
    def consistent_ic(params):
        """yy0 = [1,0,0], yp0 = [-k1, k1, 0] (examples/roberts.rs:67-68 with k1 = 0.04)"""
        p = np.asarray(params)
        batch = p.shape[1]
        yy0 = np.zeros((3, batch))
        yy0[0] = 1.0
        yp0 = np.zeros((3, batch))
        yp0[0] = -p[0]
        yp0[1] = p[0]
        return yy0, yp0


class Lorenz63(EnsembleProblem):
    """Lorenz-63 posed as F = y' - f(y) (tests/lorenz63.rs:11-52); params rows (p, r, b), defaults
    (10, 28, 8/3)."""
    kind, n, nparams = PROBLEM_LORENZ63, 3, 3
    DEFAULT = (10.0, 28.0, 8.0 / 3.0)

    @staticmethod
    def consistent_ic(params, y0=(1.0, 1.0, 1.0)):
        p = np.asarray(params)
        batch = p.shape[1]
        yy0 = np.repeat(np.array(y0, dtype=np.float64)[:, None], batch, axis=1).copy()
        s, r, b = p[0], p[1], p[2]
        yp0 = np.stack([s * (yy0[1] - yy0[0]), yy0[0] * (r - yy0[2]) - yy0[1], yy0[0] * yy0[1] - b * yy0[2]])
        return yy0, np.ascontiguousarray(yp0)


class Heat2D(EnsembleProblem):
    """2-D heat equation as a DAE on an M x M grid (n = M*M), after SUNDIALS idaHeat2D: interior
    u' = coeff * (5-point Laplacian numerator), boundary rows algebraic u = 0.  Not in the reference
    (SURVEY.md 8d config 5); params row (coeff), nominal coeff = 1/dx^2, dx = 1/(M-1)."""
    kind, nparams = PROBLEM_HEAT2D, 1

    def __init__(self, params, M=10):
        self.M = int(M)
        self.n = self.M * self.M
        super().__init__(params)

    @staticmethod
    def nominal_coeff(M):
        dx = 1.0 / (M - 1)
        return 1.0 / (dx * dx)

    @staticmethod
    def consistent_ic(params, M=10):
        """u0 = 16 x(1-x) y(1-y); u0' = coeff * laplacian on the interior, 0 on the boundary"""
        p = np.asarray(params)
        batch = p.shape[1]
        dx = 1.0 / (M - 1)
        xs = np.arange(M) * dx
        u = 16.0 * np.outer(xs * (1.0 - xs), xs * (1.0 - xs))  # [row j (y), col i (x)]
        u[0, :] = u[-1, :] = 0.0
        u[:, 0] = u[:, -1] = 0.0
        lap = np.zeros_like(u)
        lap[1:-1, 1:-1] = (u[1:-1, :-2] + u[1:-1, 2:] - 2.0 * u[1:-1, 1:-1]) + (u[:-2, 1:-1] + u[2:, 1:-1] - 2.0 * u[1:-1, 1:-1])
        yy0 = np.repeat(u.reshape(-1, 1), batch, axis=1).copy()
        yp0 = np.ascontiguousarray(lap.reshape(-1, 1) * p[0][None, :])
        return yy0, yp0


def heat_ensemble_params(batch, M=10, seed=1234, spread=0.1, first_id=0):
    """coeff = (1/dx^2) (1 + spread*u), u ~ U(-1,1) keyed by instance id (config 5)."""
    ids = np.arange(first_id, first_id + batch, dtype=np.uint64)
    c = Heat2D.nominal_coeff(M) * (1.0 + spread * (2.0 * splitmix64_uniform(seed, ids, 0) - 1.0))
    return np.ascontiguousarray(c[None, :])


def splitmix64_uniform(seed, ids, stream):
    """counter-based U[0,1): splitmix64 of (seed, instance id, stream) -- identical parameters for any
    sharding of the ensemble."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.asarray(ids, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def roberts_ensemble_params(batch, seed=1234, spread=0.1, first_id=0):
    """k_j = k_j0 * (1 + spread*u), u ~ U(-1,1) keyed by instance id (config 2)."""
    ids = np.arange(first_id, first_id + batch, dtype=np.uint64)
    rows = [k0 * (1.0 + spread * (2.0 * splitmix64_uniform(seed, ids, j) - 1.0)) for j, k0 in enumerate(Roberts.DEFAULT)]
    return np.ascontiguousarray(np.stack(rows))


def lorenz_ensemble_params(batch, seed=1234, first_id=0):
    """r ~ U(20,35), p ~ U(8,12), b = 8/3 (1 + 0.05u) (config 3)."""
    ids = np.arange(first_id, first_id + batch, dtype=np.uint64)
    p = 8.0 + 4.0 * splitmix64_uniform(seed, ids, 0)
    r = 20.0 + 15.0 * splitmix64_uniform(seed, ids, 1)
    b = (8.0 / 3.0) * (1.0 + 0.05 * (2.0 * splitmix64_uniform(seed, ids, 2) - 1.0))
    return np.ascontiguousarray(np.stack([p, r, b]))
