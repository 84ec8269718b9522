"""Tolerance control, mirroring src/tol_control.rs: ewt_i = 1 / (rtol*|y_i| + atol_i)."""
import numpy as np


class TolControlSV:
    """scalar relative tolerance, vector absolute tolerance (src/tol_control.rs:51-83)"""

    def __init__(self, rtol, atol):
        self.ida_rtol = float(rtol)
        self.ida_atol = np.ascontiguousarray(atol, dtype=np.float64)

    @classmethod
    def new(cls, rtol, atol):
        return cls(rtol, atol)

    def atol_vector(self, n):
        if self.ida_atol.shape != (n,):
            raise ValueError("atol must have %d entries" % n)
        return self.ida_atol


class TolControlSS(TolControlSV):
    """scalar relative and absolute tolerances (src/tol_control.rs:16-45)"""

    def __init__(self, rtol, atol):
        self.ida_rtol = float(rtol)
        self.ida_atol = float(atol)

    def atol_vector(self, n):
        return np.full(n, self.ida_atol, dtype=np.float64)
