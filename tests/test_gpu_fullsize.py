"""GPU checks at BASELINE.json's FULL sizes.  The headline Roberts run is compared with the (libm-pow) oracle
EXHAUSTIVELY -- all 2^20 instances, final states and every counter, bit for bit (the multi-threaded oracle integrates
them in seconds).  Lorenz (4 x 2^20) and heat (n = 100) combine size-independent properties the domain offers -- every
instance succeeds, the Roberts mass balance y1 + y2 + y3 = 1 (the algebraic third equation,
src/sample_problems/roberts.rs:58-61), the heat problem's algebraic boundary rows and maximum principle, P A = L U,
A x = b, run-to-run determinism -- with bit-exact parity against the oracle on a large random sample of instances
drawn from all over the full-size run, and split == whole."""
import os
import math

import numpy as np
import pytest

from helpers import (LORENZ_ATOL, LORENZ_RTOL, LORENZ_TOUTS, ROBERTS_ATOL, ROBERTS_RTOL, ROBERTS_TOUTS, assert_bit_equal,
                     parity_oracle)

pytestmark = pytest.mark.gpu


def _solve(problem, yy0, yp0, rtol, atol, touts):
    import torch
    import rust_ida_b200 as r
    ida = r.Ida(problem, torch.from_numpy(yy0).cuda(), torch.from_numpy(yp0).cuda(), r.TolControlSV(rtol, atol))
    status, _ = ida.solve_ladder(touts, want_traj=False)
    out = {"status": status.cpu().numpy(), "yy": ida.get_yy().cpu().numpy(), "yp": ida.get_yp().cpu().numpy(),
           "counters": {k: v.cpu().numpy() for k, v in ida.counters().items()}}
    ida.close()
    return out


def test_roberts_one_million_instances():
    """BASELINE.json configs[1]: 2^20 rate-perturbed Roberts instances to t = 4e10"""
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200.problems import roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 1 << 20
    p = roberts_ensemble_params(B, seed=1234)
    yy0, yp0 = r.Roberts.consistent_ic(p)
    g = _solve(r.Roberts(torch.from_numpy(p).cuda()), yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS)
    assert (g["status"] == 0).all()
    nst = g["counters"]["nst"]
    assert 300 < nst.min() and nst.max() < 470 and 360 < nst.mean() < 375
    # mass balance: the third residual is y1 + y2 + y3 - 1, solved by Newton to the WRMS tolerance every step
    assert np.abs(g["yy"].sum(axis=0) - 1.0).max() < 1e-6
    assert (g["yy"] > -1e-6).all() and (g["yy"][0] < 1e-5).all()  # y1 has decayed, nothing went negative beyond atol
    # bit parity with the libm-pow oracle on EVERY instance: final y, y' and all counters
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, nthreads=os.cpu_count() or 8)
    assert (ref["status"] == 0).all()
    assert_bit_equal(g["yy"], ref["yy"], "yy of all 2^20 instances")
    assert_bit_equal(g["yp"], ref["yp"], "yp of all 2^20 instances")
    for k in ("nst", "nre", "nje", "nni", "netf", "ncfn", "nsetups", "nattempts"):
        assert np.array_equal(g["counters"][k], ref["counters"][k]), k
    # determinism: a second full-size run gives the same bits
    g2 = _solve(r.Roberts(torch.from_numpy(p).cuda()), yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS)
    assert_bit_equal(g2["yy"], g["yy"], "second run")
    assert np.array_equal(g2["counters"]["nattempts"], g["counters"]["nattempts"])
    # split == whole: a shard integrated on its own equals its slice of the full ensemble (what multi-GPU sharding relies on)
    lo, hi = 3 * (B // 8), 3 * (B // 8) + 50_000
    gs = _solve(r.Roberts(torch.from_numpy(np.ascontiguousarray(p[:, lo:hi])).cuda()), np.ascontiguousarray(yy0[:, lo:hi]),
                np.ascontiguousarray(yp0[:, lo:hi]), ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS)
    assert_bit_equal(gs["yy"], g["yy"][:, lo:hi], "shard vs whole")


def test_lorenz_four_million_instances():
    """BASELINE.json configs[2]: 4 x 2^20 Lorenz-63 instances (parameter sweep) to t = 5"""
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200.problems import lorenz_ensemble_params
    ob, _ = parity_oracle()
    B = 4 << 20
    p = lorenz_ensemble_params(B, seed=1234)
    yy0, yp0 = r.Lorenz63.consistent_ic(p)
    g = _solve(r.Lorenz63(torch.from_numpy(p).cuda()), yy0, yp0, LORENZ_RTOL, LORENZ_ATOL, LORENZ_TOUTS)
    assert (g["status"] == 0).all()
    assert np.isfinite(g["yy"]).all() and np.abs(g["yy"]).max() < 100.0  # on the attractor's scale
    ids = np.unique(np.concatenate([[0, B - 1], np.random.default_rng(6).integers(0, B, 65534)]))
    ref = ob.ensemble_solve(ob.PK_LORENZ63, 3, np.ascontiguousarray(p[:, ids]), np.ascontiguousarray(yy0[:, ids]),
                            np.ascontiguousarray(yp0[:, ids]), LORENZ_RTOL, LORENZ_ATOL, LORENZ_TOUTS, nthreads=os.cpu_count() or 8)
    assert_bit_equal(g["yy"][:, ids], ref["yy"], "yy of sampled instances")
    for k in ("nst", "nni", "nje", "netf", "nattempts"):
        assert np.array_equal(g["counters"][k][ids], ref["counters"][k]), k


def test_heat_sixty_four_thousand_instances():
    """BASELINE.json configs[4]: 65,536 heat-2D instances (dense n = 100), CTA-per-system kernel"""
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200.problems import heat_ensemble_params
    ob, _ = parity_oracle()
    B, M = 65536, 10
    n = M * M
    p = heat_ensemble_params(B, M=M, seed=1234)
    yy0, yp0 = r.Heat2D.consistent_ic(p, M)
    touts = 0.01 * 2.0 ** np.arange(11)
    rtol, atol = 1e-6, np.full(n, 1e-8)
    g = _solve(r.Heat2D(torch.from_numpy(p).cuda(), M=M), yy0, yp0, rtol, atol, touts)
    assert (g["status"] == 0).all()
    u = g["yy"].reshape(M, M, B)
    boundary = np.concatenate([u[0].ravel(), u[-1].ravel(), u[:, 0].ravel(), u[:, -1].ravel()])
    assert np.abs(boundary).max() < 1e-12            # algebraic rows u = 0
    assert u.max() < yy0.max() and u.min() > -1e-6   # maximum principle: the solution only decays
    ids = np.unique(np.concatenate([[0, B - 1], np.random.default_rng(7).integers(0, B, 22)]))
    ref = ob.ensemble_solve(ob.PK_HEAT2D, n, p[:, ids], yy0[:, ids], yp0[:, ids], rtol, atol, touts, nthreads=8)
    assert_bit_equal(g["yy"][:, ids], ref["yy"], "yy of sampled instances")
    for k in ("nst", "nni", "nje", "netf", "nattempts"):
        assert np.array_equal(g["counters"][k][ids], ref["counters"][k]), k


@pytest.mark.parametrize("n", [3, 16, 32, 64, 100, 128, 256])
def test_lu_at_benchmark_batch_sizes(n):
    """BASELINE.json configs[3]: batch = ceil(9 * 2^20 / n^2) matrices of the reference's generator.  P A = L U and
    A x = b on every matrix, pivots legal, and bit parity with the oracle on a sample."""
    import torch
    import rust_ida_b200 as r
    from oracle import binding as ob
    batch = math.ceil(9 * (1 << 20) / (n * n))
    rng = np.random.default_rng(1234 + n)
    a = rng.uniform(0.0, 1.0 / n, size=(batch, n, n))
    a += np.eye(n)[::-1][None]
    b = rng.uniform(0.0, 1.0, size=(batch, n))
    layout = "soa" if n <= 8 else "aos"
    if layout == "soa":
        A = torch.from_numpy(np.ascontiguousarray(a.transpose(1, 2, 0))).cuda()
        X = torch.from_numpy(np.ascontiguousarray(b.T)).cuda()
    else:
        A = torch.from_numpy(np.ascontiguousarray(a.transpose(0, 2, 1))).cuda()
        X = torch.from_numpy(b.copy()).cuda()
    ls = r.Dense.new(n, batch, layout=layout)
    ls.setup(A)
    ls.solve(A, X, X.clone(), 0.0)
    torch.cuda.synchronize()
    if layout == "soa":
        lu, piv, x = A.cpu().numpy().transpose(2, 0, 1), ls.pivots.cpu().numpy().T, X.cpu().numpy().T
    else:
        lu, piv, x = A.cpu().numpy().transpose(0, 2, 1), ls.pivots.cpu().numpy(), X.cpu().numpy()
    assert (ls.info.cpu().numpy() == 0).all()
    k = np.arange(n)
    assert (piv >= k[None, :]).all() and (piv < n).all()
    # A x = b on every system
    res = np.abs(np.einsum("bij,bj->bi", a, x) - b).max()
    assert res < 1e-10, res
    # P A = L U on a sample, and bit parity with the oracle on the same sample
    ids = np.unique(np.concatenate([[0, batch - 1], rng.integers(0, batch, 30)]))
    for m in ids:
        L = np.tril(lu[m], -1) + np.eye(n)
        U = np.triu(lu[m])
        pa = a[m].copy()
        for kk in range(n):
            pa[[kk, piv[m, kk]]] = pa[[piv[m, kk], kk]]
        assert np.abs(L @ U - pa).max() < 1e-12
        lu_ref, piv_ref, info = ob.getrf(a[m])
        assert info == 0 and np.array_equal(piv_ref, piv[m])
        assert_bit_equal(lu[m], lu_ref, "LU of matrix %d" % m)
        assert_bit_equal(x[m], ob.getrs(lu_ref, piv_ref, b[m]), "x of matrix %d" % m)
