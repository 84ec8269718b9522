"""NLSolver level of the boundary (b200ida_newton_batched, rust_ida_b200.nonlinear.Newton) against the reference's own
Newton known answer (crates/nonlinear/src/newton.rs:306-343) and against the oracle's Newton on perturbed batches."""
import numpy as np
import pytest

from helpers import assert_bit_equal, parity_oracle

pytestmark = pytest.mark.gpu

Y_EXP = np.array([0.785196933062355226, 0.496611392944656396, 0.369922830745872357])   # newton.rs:309-313
EXPECTED_ERR = np.array([-0.00578453, 1.0143e-08, 1.47767e-08])                       # newton.rs:334


def test_reference_known_answer():
    """test_newton of the reference: y0 = (.5, .5, .5), w = 1, tol 1e-2, Newton::new(10), call_lsetup = true;
    assert_relative_eq!(y - y_exp, expected_err, max_relative = 1e-5)"""
    import rust_ida_b200 as r
    y0 = np.full((3, 1), 0.5)
    w = np.ones((3, 1))
    y = np.zeros((3, 1))
    newton = r.Newton.new(10)
    st = newton.solve(r.Quad3Problem(), y0, y, w, 1e-2, True)
    assert st[0] == 0
    err = y[:, 0] - Y_EXP
    assert np.all(np.abs(err - EXPECTED_ERR) <= 1e-5 * np.maximum(np.abs(err), np.abs(EXPECTED_ERR)))
    ob, _ = parity_oracle()
    rr, yy, ni, nc = ob.newton_quad3(y0[:, 0], w[:, 0], 1e-2, 10)
    assert (rr, ni, nc) == (0, 6, 0) and newton.get_num_iters()[0] == ni and newton.get_num_conv_fails()[0] == nc
    assert_bit_equal(y[:, 0], yy, "iterate vs oracle")


@pytest.mark.parametrize("device_buffers", [False, True])
def test_perturbed_batch_matches_the_oracle(device_buffers):
    """20,000 systems: perturbed initial guesses and weights, a tolerance tight enough that some systems need the
    retry on a fresh Jacobian or run out of iterations -> statuses, counters and iterates equal the oracle's bit for bit"""
    import torch
    import rust_ida_b200 as r
    ob, _ = parity_oracle()
    rng = np.random.default_rng(3)
    B = 20000
    y0 = 0.5 + 0.45 * rng.uniform(-1, 1, (3, B))
    y0[:, :50] *= 4.0  # far-off guesses: slow convergence / recover
    w = 10.0 ** rng.uniform(-1, 3, (3, B))
    for tol, maxiters, lsetup in ((1e-2, 10, True), (1e-9, 4, True), (1e-6, 3, False)):
        y = np.zeros((3, B))
        newton = r.Newton.new(maxiters)
        if device_buffers:
            a, bb, c = (torch.from_numpy(v).cuda() for v in (y0, y, w))
            st = newton.solve(r.Quad3Problem(), a, bb, c, tol, lsetup).cpu().numpy()
            y = bb.cpu().numpy()
            ni, nc = newton.get_num_iters().cpu().numpy(), newton.get_num_conv_fails().cpu().numpy()
        else:
            st = newton.solve(r.Quad3Problem(), y0, y, w, tol, lsetup)
            ni, nc = newton.get_num_iters(), newton.get_num_conv_fails()
        ref_y = np.zeros((3, B))
        ref = np.zeros((3, B), dtype=np.int64)
        for b in range(B if not lsetup else 4000):  # the scalar oracle call through ctypes: a sample is enough for two of the legs
            rr, yy, n1, n2 = ob.newton_quad3(y0[:, b].copy(), w[:, b].copy(), tol, maxiters, call_lsetup=lsetup)
            ref_y[:, b] = yy
            ref[:, b] = (rr, n1, n2)
        m = B if not lsetup else 4000
        assert np.array_equal(st[:m], ref[0, :m]) and np.array_equal(ni[:m], ref[1, :m]) and np.array_equal(nc[:m], ref[2, :m])
        assert_bit_equal(y[:, :m], ref_y[:, :m], "Newton iterates")
    assert (st != 0).any()  # the no-lsetup leg solves with an all-zero (never factored) matrix: division by zero -> no convergence


def test_singular_jacobian_is_reported_per_system():
    import rust_ida_b200 as r
    y0 = np.array([[0.0, 0.5], [0.5, 0.5], [0.5, 0.5]])  # x = 0: first Jacobian column is zero -> LUFactFail { col: 1 }
    y = np.zeros((3, 2))
    st = r.Newton.new(10).solve(r.Quad3Problem(), y0, y, np.ones((3, 2)), 1e-2, True)
    assert st[0] == -1 and st[1] == 0
