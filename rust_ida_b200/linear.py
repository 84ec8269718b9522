"""Batched mirror of the reference's linear-solver plugin interface (crates/linear).

`Dense` keeps the names, argument meaning and error behaviour of `linear::Dense<D>` /
`LSolver<T, D>` (crates/linear/src/traits.rs:27-91, crates/linear/src/dense.rs:13-63) but factors
and solves a whole ensemble of n x n systems per call on the GPU through the C ABI
(b200ida_getrf_batched / b200ida_getrs_batched).  Pivots are kept inside the solver, exactly like the
reference's `Dense { pivots }`.

Layouts (see include/b200ida.h):  layout="soa": matrices (n, n, batch) with A[i, j, b] = a_ij, vectors
(n, batch);  layout="aos": matrices (batch, n, n) in nalgebra (column-major) storage, i.e.
A[b, j, i] = a_ij, vectors (batch, n).  numpy arrays = host buffers, torch CUDA tensors = device.
"""
import enum

import numpy as np

from . import _buffers as B
from . import _lib


class LSolverType(enum.Enum):  # crates/linear/src/lib.rs:15-20
    Direct = 0
    Iterative = 1
    MatrixIterative = 2


class LUFactFail(Exception):
    """linear::Error::LUFactFail { col } (crates/linear/src/lib.rs:9-13); `col` is 1-based.
    `index` is the first failing system of the batch, `info` the whole per-system array."""

    def __init__(self, col, index, info):
        super().__init__("A singular matrix was encountered during a LU factorization (col %d) [system %d]" % (col, index))
        self.col, self.index, self.info = col, index, info


def _layout_code(layout):
    return {"soa": _lib.LAYOUT_SOA, "aos": _lib.LAYOUT_AOS}[layout]


class Dense:
    """`impl LSolver<f64, D> for Dense<D>` over a batch (crates/linear/src/dense.rs:22-63)."""

    def __init__(self, n, batch, layout="soa", raise_on_singular=True, tensor_core_update=False, team_panel=False):
        """tensor_core_update: TOLERANCE mode (B200IDA_LU_DMMA) -- the trailing update of the blocked LU on the FP64
        tensor cores for AoS matrices with even n >= 115; factors agree with dense_get_rf to rounding, not bit for bit.
        team_panel: B200IDA_LU_TEAM_PANEL -- the CTA-team IDA kernel's panel / update schedule for AoS, 33 <= n <= 160
        (bit-identical results; for tests and measurements)."""
        self.n, self.batch, self.layout = int(n), int(batch), layout
        self.raise_on_singular = raise_on_singular
        self.tensor_core_update = bool(tensor_core_update)
        self.team_panel = bool(team_panel)
        self.pivots = None  # int64, (n, batch) soa / (batch, n) aos -- `Dense { pivots }`
        self.info = None    # int32 (batch,): 0 or 1-based zero-pivot column

    @classmethod
    def new(cls, n, batch, layout="soa", **kw):  # LSolver::new, traits.rs:32
        return cls(n, batch, layout, **kw)

    def get_type(self):  # traits.rs:34, dense.rs:34-36
        return LSolverType.Direct

    def set_scaling_vectors(self, _s1, _s2):  # traits.rs:46 (no-op for a direct solver)
        return None

    def _mat_shape(self):
        return (self.n, self.n, self.batch) if self.layout == "soa" else (self.batch, self.n, self.n)

    def _vec_shape(self):
        return (self.n, self.batch) if self.layout == "soa" else (self.batch, self.n)

    def setup(self, mat_a):
        """LSolver::setup (traits.rs:52-54): in-place LU of every matrix; pivots kept in the solver.
        Raises LUFactFail for the first singular system (the reference returns Err(LUFactFail{col}))."""
        _lib.require_device()
        if tuple(mat_a.shape) != self._mat_shape():
            raise ValueError("mat_a must have shape %r for layout %s" % (self._mat_shape(), self.layout))
        pa, mem = B.describe(mat_a, np.float64, "mat_a")
        if self.pivots is None or B.is_torch(self.pivots) != B.is_torch(mat_a):
            self.pivots = B.empty_like_space(mat_a, self._vec_shape(), np.int64)
            self.info = B.empty_like_space(mat_a, (self.batch,), np.int32)
        pp, _ = B.describe(self.pivots, np.int64, "pivots")
        pi, _ = B.describe(self.info, np.int32, "info")
        ctx, dev = B.tensors_device(mat_a, self.pivots, self.info)
        with ctx:
            code = (_layout_code(self.layout) | (_lib.LU_DMMA if self.tensor_core_update else 0) |
                    (_lib.LU_TEAM_PANEL if self.team_panel else 0))
            _lib.check(_lib.lib().b200ida_getrf_batched(pa, pp, pi, self.n, self.batch, code, mem, B.current_stream(mem, dev)))
        if self.raise_on_singular:
            info = self.info.cpu().numpy() if B.is_torch(self.info) else self.info
            bad = np.nonzero(info)[0]
            if bad.size:
                raise LUFactFail(int(info[bad[0]]), int(bad[0]), info)

    def solve(self, mat_a, x, b, tol=0.0):
        """LSolver::solve (traits.rs:70-80): copy b into x (dense.rs:58-59), then getrs in place on x."""
        _lib.require_device()
        if self.pivots is None:
            raise RuntimeError("Dense.solve called before setup")
        if tuple(x.shape) != self._vec_shape() or tuple(b.shape) != self._vec_shape():
            raise ValueError("x and b must have shape %r" % (self._vec_shape(),))
        if B.is_torch(x):
            x.copy_(b)
        else:
            np.copyto(x, b)
        pa, m1 = B.describe(mat_a, np.float64, "mat_a")
        px, m2 = B.describe(x, np.float64, "x")
        pp, m3 = B.describe(self.pivots, np.int64, "pivots")
        mem = B.common_mem(m1, m2, m3)
        ctx, dev = B.tensors_device(mat_a, x, self.pivots)
        with ctx:
            _lib.check(_lib.lib().b200ida_getrs_batched(pa, pp, px, self.n, self.batch, _layout_code(self.layout), mem,
                                                        B.current_stream(mem, dev)))

    def num_iters(self):  # traits.rs:83
        return 0

    def res_norm(self):  # traits.rs:88
        return 0.0


def norm_wrms(x, w, layout="soa"):
    """NormRms::norm_wrms / NormWRMS::norm_wrms for every system of a batch (src/norm_rms.rs:31-38,
    crates/nonlinear/src/norm_wrms.rs:24-33).  x, w: (n, batch) soa or (batch, n) aos."""
    _lib.require_device()
    n, batch = (x.shape if layout == "soa" else x.shape[::-1])
    px, m1 = B.describe(x, np.float64, "x")
    pw, m2 = B.describe(w, np.float64, "w")
    mem = B.common_mem(m1, m2)
    out = B.empty_like_space(x, (batch,), np.float64)
    po, _ = B.describe(out, np.float64, "out")
    ctx, dev = B.tensors_device(x, w)
    with ctx:
        _lib.check(_lib.lib().b200ida_wrms_batched(px, pw, po, int(n), int(batch), _layout_code(layout), mem,
                                                   B.current_stream(mem, dev)))
    return out


def norm_wrms_masked(x, w, id_, layout="soa"):
    """NormRmsMasked::norm_wrms_masked for every system of a batch (src/norm_rms.rs:49-57): components with
    id == False are left out of the sum, the divisor stays n.  x, w: (n, batch) soa or (batch, n) aos; id_: (n,) bools,
    one mask for the whole batch (`ida_id`, src/lib.rs:122)."""
    _lib.require_device()
    n, batch = (x.shape if layout == "soa" else x.shape[::-1])
    px, m1 = B.describe(x, np.float64, "x")
    pw, m2 = B.describe(w, np.float64, "w")
    mem = B.common_mem(m1, m2)
    mask = np.ascontiguousarray(np.asarray(id_).reshape(-1) != 0, dtype=np.uint8)
    if mask.size != n:
        raise ValueError("id has %d entries, the systems have %d components" % (mask.size, n))
    if mem == _lib.MEM_DEVICE:
        import torch
        mask_t = torch.from_numpy(mask).to(x.device)
        pm = mask_t.data_ptr()
    else:
        pm = mask.ctypes.data
    out = B.empty_like_space(x, (batch,), np.float64)
    po, _ = B.describe(out, np.float64, "out")
    ctx, dev = B.tensors_device(x, w)
    with ctx:
        _lib.check(_lib.lib().b200ida_wrms_masked_batched(px, pw, pm, po, int(n), int(batch), _layout_code(layout), mem,
                                                          B.current_stream(mem, dev)))
    return out
