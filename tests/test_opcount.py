"""The oracle's algorithmic operation counter (oracle/ida_oracle.cpp, -DORACLE_COUNT_OPS) -- the work unit of bench.py's
FP64 roofline.  Pinned on the reference's own example: the default Roberts instance of examples/roberts.rs."""
import numpy as np

from helpers import ROBERTS_ATOL, ROBERTS_RTOL, ROBERTS_TOUTS


def test_default_roberts_operation_counts():
    from oracle import binding as ob
    from rust_ida_b200.problems import Roberts
    p = np.array([[0.04], [1.0e4], [3.0e7]])
    yy0, yp0 = Roberts.consistent_ic(p)
    ops, natt = ob.ensemble_ops(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS)
    add, mul, div, sqrt, pw = (int(v) for v in ops[0])
    assert natt[0] == 377                     # data_trace frames of the reference author's notebook
    # SURVEY.md section 6 (independent probe): 2,030 WRMS norms (one sqrt each) and 528 pow calls per instance
    assert sqrt == 2030 and pw == 528
    # the counter's own totals, pinned so that a change of what is counted shows up in review
    assert (add, mul, div) == (47848, 40932, 12266)
    assert abs((add + mul + div + sqrt + pw) / natt[0] - 274.81) < 0.01


def test_counting_build_integrates_identically():
    """the -DORACLE_COUNT_OPS build is the same restatement: same attempts as the production oracle"""
    from oracle import binding as ob
    from rust_ida_b200.problems import Roberts, roberts_ensemble_params
    p = roberts_ensemble_params(64, seed=5)
    yy0, yp0 = Roberts.consistent_ic(p)
    _, natt = ob.ensemble_ops(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, nthreads=4)
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, nthreads=4)
    assert np.array_equal(natt, ref["counters"]["nattempts"])
