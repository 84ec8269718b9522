"""rust_ida_b200 -- B200-native batched IDA engine: the data-parallel hot path of jondo2010/rust-ida
(per-step BDF/Newton loop over ensembles of independent DAE instances) as hand-written sm_100a CUDA
behind a C ABI (include/b200ida.h), with a host-side mirror of the reference's interfaces:

    linear.Dense / LSolverType / LUFactFail   <- crates/linear (LSolver trait, Dense LU)
    linear.norm_wrms, norm_wrms_masked        <- src/norm_rms.rs, crates/nonlinear/src/norm_wrms.rs
    tol_control.TolControlSV / TolControlSS   <- src/tol_control.rs
    problems.Roberts / Lorenz63               <- src/sample_problems, tests/lorenz63.rs
    Ida / IdaTask / IdaSolveStatus            <- src/lib.rs, src/impl_solve.rs, src/ida_io.rs

No CPU fallback: compute calls raise if libb200ida.so or a CUDA device is missing.
"""
from . import _lib
from ._lib import B200IdaError, build, device_count
from .ida import Ida, IdaSolveStatus, IdaTask, set_default_reorder
from .linear import Dense, LSolverType, LUFactFail, norm_wrms, norm_wrms_masked
from .nonlinear import Newton, Quad3Problem
from .problems import Heat2D, Lorenz63, Roberts
from .tol_control import TolControlSS, TolControlSV

__all__ = ["Ida", "IdaTask", "IdaSolveStatus", "Dense", "LSolverType", "LUFactFail", "norm_wrms", "norm_wrms_masked", "Roberts",
           "Lorenz63", "Heat2D", "TolControlSV", "TolControlSS", "B200IdaError", "build", "device_count", "set_default_reorder"]
