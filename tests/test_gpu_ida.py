"""GPU parity: the ensemble IDA engine through the C ABI vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): per-instance step / Newton / Jacobian-evaluation counts identical, states
within rtol 1e-10 + atol 1e-14.  With un-fused arithmetic and the glibc-compatible pow the engine is in
fact bit-identical to the oracle, so the tests assert bit equality and the tolerance as a fallback bar.
"""
import numpy as np
import pytest

from helpers import (LORENZ_ATOL, LORENZ_RTOL, LORENZ_TOUTS, ROBERTS_ATOL, ROBERTS_RTOL, ROBERTS_TOUTS,
                     assert_bit_equal, parity_oracle)

pytestmark = pytest.mark.gpu

COUNTERS = ("nst", "nre", "nje", "nni", "netf", "ncfn", "nsetups", "nattempts")


def run_engine(problem_cls, params, yy0, yp0, rtol, atol, touts, ladder, device_buffers=False, kernel="auto",
               problem_kw=None, id_mask=None):
    import torch
    import rust_ida_b200 as r
    if device_buffers:
        params, yy0, yp0 = (torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in (params, yy0, yp0))
    ida = r.Ida(problem_cls(params, **(problem_kw or {})), yy0, yp0, r.TolControlSV(rtol, atol))
    ida.set_kernel(kernel)
    if id_mask is not None:
        ida.set_id(id_mask, suppressalg=True)
    if ladder:
        status, traj = ida.solve_ladder(touts)
    else:
        sts, trs = [], []
        for t in touts:
            st, _ = ida.solve(t)
            sts.append(st)
            trs.append(ida.get_yy())
        status = np.stack([np.asarray(s.cpu() if device_buffers else s) for s in sts])
        traj = np.stack([np.asarray(t.cpu() if device_buffers else t) for t in trs])
    to_np = (lambda v: v.cpu().numpy()) if device_buffers else (lambda v: np.asarray(v))
    out = {"status": to_np(status) if ladder else status, "traj": to_np(traj) if ladder else traj,
           "yy": to_np(ida.get_yy()), "yp": to_np(ida.get_yp()),
           "counters": {k: to_np(v) for k, v in ida.counters().items()},
           "kused": to_np(ida.get_last_order()), "kk": to_np(ida.get_current_order()),
           "hused": to_np(ida.get_last_step()), "hh": to_np(ida.get_current_step()),
           "tn": to_np(ida.get_current_time()), "h0u": to_np(ida.get_actual_init_step())}
    ida.close()
    return out


def compare(gpu, ref, n_expected_fail=0):
    for k in COUNTERS:
        assert np.array_equal(gpu["counters"][k], ref["counters"][k]), "counter %s differs" % k
    assert np.array_equal(gpu["status"], ref["status"])
    assert np.array_equal(gpu["kused"], ref["kused"]) and np.array_equal(gpu["kk"], ref["kk"])
    # the stated bar ...
    np.testing.assert_allclose(gpu["yy"], ref["yy"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(gpu["traj"], ref["yy_traj"], rtol=1e-10, atol=1e-14)
    # ... and what the engine actually delivers: bit-identical states, step sizes and times
    assert_bit_equal(gpu["yy"], ref["yy"], "yy")
    assert_bit_equal(gpu["yp"], ref["yp"], "yp")
    assert_bit_equal(gpu["traj"], ref["yy_traj"], "yy_traj")
    assert_bit_equal(gpu["hused"], ref["hused"], "hused")
    assert_bit_equal(gpu["hh"], ref["hh"], "hh")
    assert_bit_equal(gpu["tn"], ref["tn"], "tn")


def test_roberts_single_instance_matches_reference_run():
    """config 1 (examples/roberts.rs): nst/attempt/nre/nje/netf counts and the final state"""
    from rust_ida_b200 import Roberts
    p = np.array(Roberts.DEFAULT).reshape(3, 1)
    yy0, yp0 = Roberts.consistent_ic(p)
    assert yp0[:, 0].tolist() == [-0.04, 0.04, 0.0]
    g = run_engine(Roberts, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, ladder=False)
    c = {k: int(v[0]) for k, v in g["counters"].items()}
    assert (c["nst"], c["nattempts"], c["nre"], c["nni"], c["nje"], c["netf"], c["ncfn"]) == (362, 377, 537, 537, 60, 15, 0)
    assert g["yy"][:, 0].tolist() == [4.86408986935633547e-08, 1.94563604063205543e-13, 9.99999951358906736e-01]
    assert (g["status"] == 0).all()
    assert int(g["kused"][0]) == 2 and abs(g["hused"][0] / 7.54805e9 - 1) < 1e-5


@pytest.mark.parametrize("ladder", [False, True])
@pytest.mark.parametrize("device_buffers", [False, True])
def test_roberts_ensemble_parity(ladder, device_buffers):
    """config 2 at test size: 4,224 rate-perturbed instances (33 CTAs incl. a ragged tail)"""
    from rust_ida_b200 import Roberts
    from rust_ida_b200.problems import roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 4224 - 37
    p = roberts_ensemble_params(B, seed=1234)
    yy0, yp0 = Roberts.consistent_ic(p)
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, nthreads=8,
                            want_traj=True)
    g = run_engine(Roberts, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, ladder, device_buffers)
    compare(g, ref)
    nst = g["counters"]["nst"]
    assert 300 < nst.min() and nst.max() < 450 and nst.std() > 3  # a genuinely heterogeneous ensemble
    assert (g["counters"]["ncfn"] > 0).any()  # the Newton-recover path is exercised


def test_lorenz_ensemble_parity():
    """config 3 at test size"""
    from rust_ida_b200 import Lorenz63
    from rust_ida_b200.problems import lorenz_ensemble_params
    ob, _ = parity_oracle()
    B = 2048 + 5
    p = lorenz_ensemble_params(B, seed=1234)
    yy0, yp0 = Lorenz63.consistent_ic(p)
    ref = ob.ensemble_solve(ob.PK_LORENZ63, 3, p, yy0, yp0, LORENZ_RTOL, LORENZ_ATOL, LORENZ_TOUTS, nthreads=8,
                            want_traj=True)
    g = run_engine(Lorenz63, p, yy0, yp0, LORENZ_RTOL, LORENZ_ATOL, LORENZ_TOUTS, ladder=True)
    compare(g, ref)
    assert (g["status"] == 0).all()


def test_mxstep_is_per_call_and_resumable():
    """mxstep = 500 per solve() call (src/impl_solve.rs:249-260): Lorenz needs 613 steps to t = 5, so a single
    solve(5.0) returns TooMuchWork after exactly 500 steps, and calling again continues from there -- same as
    the oracle."""
    from rust_ida_b200 import IdaSolveStatus, Lorenz63
    ob, _ = parity_oracle()
    p = np.array(Lorenz63.DEFAULT).reshape(3, 1).repeat(3, axis=1).copy()
    yy0, yp0 = Lorenz63.consistent_ic(p)
    touts = np.array([5.0, 5.0, 6.0])
    ref = ob.ensemble_solve(ob.PK_LORENZ63, 3, p, yy0, yp0, LORENZ_RTOL, LORENZ_ATOL, touts, want_traj=True)
    g = run_engine(Lorenz63, p, yy0, yp0, LORENZ_RTOL, LORENZ_ATOL, touts, ladder=True)
    assert g["status"][:, 0].tolist() == [IdaSolveStatus.TooMuchWork, 0, 0] and ref["status"][:, 0].tolist() == [-1, 0, 0]
    compare(g, ref)
    # the same through separate solve() calls
    g2 = run_engine(Lorenz63, p, yy0, yp0, LORENZ_RTOL, LORENZ_ATOL, touts, ladder=False)
    compare(g2, ref)


def test_error_statuses_match_oracle():
    """ill-posed inputs: tout == t0 (IllegalInput), absurd tolerance (TooMuchAccuracy), singular
    Jacobian (k = 0 Roberts => the reference panics; engine and oracle report LinearSetupFail)."""
    from rust_ida_b200 import IdaSolveStatus, Roberts
    ob, _ = parity_oracle()
    p = np.array(Roberts.DEFAULT).reshape(3, 1).repeat(2, axis=1).copy()
    yy0, yp0 = Roberts.consistent_ic(p)
    for rtol, atol, touts in ((1e-4, ROBERTS_ATOL, np.array([0.0, 0.4])), (1e-20, np.full(3, 1e-30), np.array([0.4]))):
        ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, rtol, atol, touts, want_traj=True)
        g = run_engine(Roberts, p, yy0, yp0, rtol, atol, touts, ladder=True)
        assert np.array_equal(g["status"], ref["status"])
        assert g["status"][0, 0] in (IdaSolveStatus.IllegalInput, IdaSolveStatus.TooMuchAccuracy)
        assert_bit_equal(g["yy"], ref["yy"], "yy after error")


def test_split_ensemble_equals_whole():
    """sharding property used by the multi-GPU path: integrating two halves separately gives exactly the
    per-instance results of integrating the whole ensemble (instances are independent)."""
    from rust_ida_b200 import Roberts
    from rust_ida_b200.problems import roberts_ensemble_params
    B = 1000
    p = roberts_ensemble_params(B, seed=99)
    yy0, yp0 = Roberts.consistent_ic(p)
    whole = run_engine(Roberts, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS[:6], ladder=True)
    parts = []
    for lo, hi in ((0, 333), (333, 1000)):
        pp = np.ascontiguousarray(p[:, lo:hi])
        y, yp = Roberts.consistent_ic(pp)
        parts.append(run_engine(Roberts, pp, y, yp, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS[:6], ladder=True))
    assert_bit_equal(np.concatenate([q["yy"] for q in parts], axis=1), whole["yy"], "sharded yy")
    assert np.array_equal(np.concatenate([q["counters"]["nst"] for q in parts]), whole["counters"]["nst"])
    # shard parameters are keyed by instance id, so shards can generate their own slice
    assert np.array_equal(roberts_ensemble_params(667, seed=99, first_id=333), p[:, 333:])


@pytest.mark.parametrize("kernel", ["warp", "cta"])
def test_team_kernels_match_oracle_on_roberts(kernel):
    """warp-per-system and CTA-per-system kernels (csrc/ida_team.cuh) on the n = 3 problem: bit-identical to the
    oracle (and therefore to the thread-per-instance kernel)."""
    from rust_ida_b200 import Roberts
    from rust_ida_b200.problems import roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 777
    p = roberts_ensemble_params(B, seed=4321)
    yy0, yp0 = Roberts.consistent_ic(p)
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, nthreads=8,
                            want_traj=True)
    g = run_engine(Roberts, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, ladder=True, kernel=kernel)
    compare(g, ref)


@pytest.mark.parametrize("M,kernel,B", [(4, "warp", 300), (5, "warp", 130), (4, "cta", 64), (10, "cta", 700), (10, "auto", 41),
                                        (3, "warp", 64), (7, "cta", 30), (8, "cta", 40), (12, "cta", 12)])
def test_heat2d_parity(M, kernel, B):
    """config 5 at test size: 2-D heat DAE (n = M*M; 10x10 -> dense n = 100, CTA-per-system with the LU factors
    resident in shared memory), coefficient perturbed +-10 % per instance."""
    from rust_ida_b200 import Heat2D
    from rust_ida_b200.problems import heat_ensemble_params
    ob, _ = parity_oracle()
    n = M * M
    p = heat_ensemble_params(B, M=M, seed=1234)
    yy0, yp0 = Heat2D.consistent_ic(p, M)
    touts = 0.01 * 2.0 ** np.arange(11)
    rtol, atol = 1e-6, np.full(n, 1e-8)
    ref = ob.ensemble_solve(ob.PK_HEAT2D, n, p, yy0, yp0, rtol, atol, touts, nthreads=8, want_traj=True)
    g = run_engine(Heat2D, p, yy0, yp0, rtol, atol, touts, ladder=True, kernel=kernel, problem_kw={"M": M})
    compare(g, ref)
    assert (g["status"] == 0).all() and g["counters"]["nje"].min() > 5


def test_thread_kernel_rejects_large_n():
    import rust_ida_b200 as r
    from rust_ida_b200.problems import heat_ensemble_params
    p = heat_ensemble_params(4, M=4)
    yy0, yp0 = r.Heat2D.consistent_ic(p, 4)
    ida = r.Ida(r.Heat2D(p, M=4), yy0, yp0, r.TolControlSV(1e-6, np.full(16, 1e-8)))
    ida.set_kernel("thread")
    with pytest.raises(r.B200IdaError):
        ida.solve(0.01)


def test_cpp_host_mirror_example():
    """include/b200ida.hpp (compiled C++ mirror of Ida::new / solve / getters) through examples/roberts_ensemble:
    reproduces the reference run's 362 steps for instance 0 (exit code 0) and prints the examples/roberts.rs table."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "roberts_ensemble")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "steps 362  res evals 537  jac evals 60  nonlin iters 537  err test fails 15  conv fails 0" in out.stdout
    # the (nst, k, h) columns of the table examples/roberts.rs:87-113 prints, through the step-info getters
    table = [(29, 3, 8.80238e-02), (43, 4, 6.33772e-01), (68, 4, 3.19323e+00), (95, 4, 3.32007e+01), (126, 3, 3.14577e+02),
             (161, 5, 2.50577e+03), (202, 3, 2.63706e+04), (250, 3, 1.71865e+05), (293, 4, 2.36551e+06), (325, 4, 2.68085e+07),
             (348, 3, 7.43054e+08), (362, 2, 7.54805e+09)]
    rows = [ln.split() for ln in out.stdout.split("\n") if "status" in ln and ln[0].isdigit()]
    assert len(rows) == 12
    for row, (nst, k, h) in zip(rows, table):
        assert (int(row[4]), int(row[5])) == (nst, k) and abs(float(row[6]) / h - 1.0) < 1e-5 and row[-1] == "0"
    # get_dky(t_end, 1) agrees with the solver's yp at the output time to rounding
    dk = [ln for ln in out.stdout.split("\n") if ln.startswith("dky1")][0].split()
    nums = [float(dk[i]) for i in (1, 2, 3, 5, 7, 8, 9)]
    assert len(nums) == 7 and nums[3] == 0 and np.allclose(nums[:3], nums[4:], rtol=1e-5)


@pytest.mark.parametrize("mode", ["on", "off"])
@pytest.mark.parametrize("device_buffers", [False, True])
def test_internal_reordering_is_invisible(mode, device_buffers):
    """warp-coherence reordering (Morton order of the parameter vectors) is internal: every output stays in the
    caller's instance order and is bit-identical with it on or off, for both solve() and solve_ladder()."""
    import rust_ida_b200 as r
    from rust_ida_b200 import Roberts
    from rust_ida_b200.problems import roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 1531
    p = roberts_ensemble_params(B, seed=77)
    yy0, yp0 = Roberts.consistent_ic(p)
    touts = ROBERTS_TOUTS[:7]
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, touts, nthreads=8, want_traj=True)
    r.set_default_reorder(mode)
    try:
        for ladder in (True, False):
            g = run_engine(Roberts, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, touts, ladder, device_buffers)
            compare(g, ref)
    finally:
        r.set_default_reorder("auto")


def test_reinit_with_new_parameters_rebuilds_order():
    import rust_ida_b200 as r
    from rust_ida_b200 import Roberts
    from rust_ida_b200.problems import roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 5000  # >= 4096: AUTO reorders
    p1 = roberts_ensemble_params(B, seed=1)
    p2 = roberts_ensemble_params(B, seed=2)
    yy0, yp0 = Roberts.consistent_ic(p1)
    ida = r.Ida(Roberts(p1), yy0, yp0, r.TolControlSV(ROBERTS_RTOL, ROBERTS_ATOL))
    ida.solve_ladder(ROBERTS_TOUTS[:3], want_traj=False)
    yy0b, yp0b = Roberts.consistent_ic(p2)
    ida.reinit(yy0b, yp0b, params=p2)
    st, traj = ida.solve_ladder(ROBERTS_TOUTS[:4])
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p2, yy0b, yp0b, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS[:4], nthreads=8,
                            want_traj=True)
    assert_bit_equal(traj, ref["yy_traj"], "traj after reinit")
    assert np.array_equal(ida.get_num_steps(), ref["counters"]["nst"])
    assert_bit_equal(ida.get_last_step(), ref["hused"], "hused")


# ---- per-attempt state trace: the device side of the reference's data_trace feature (src/lib.rs:297-301, 635-639) ------
TRACE_D_FIELDS = None  # filled lazily: names of the double columns, for readable failures


def _trace_names(n):
    d = ["phi[%d][%d]" % (j, i) for j in range(6) for i in range(n)]
    for a in ("psi", "alpha", "beta", "sigma", "gamma"):
        d += ["%s[%d]" % (a, j) for j in range(6)]
    for v in ("ee", "ewt", "yy", "yp"):
        d += ["%s[%d]" % (v, i) for i in range(n)]
    d += ["hh", "hused", "rr", "cj", "cjold", "cjlast", "cjratio", "ss", "toldel", "eps_newt", "tn", "tretlast", "h0u",
          "tolsf", "tret"]
    d += ["J[%d][%d]" % (e % n, e // n) for e in range(n * n)]
    q = ["kk", "kused", "knew", "phase", "ns", "setup_done", "jcur", "nst", "nre", "nje", "nni", "netf", "ncfn", "nsetups",
         "nlscf", "natt"] + ["piv[%d]" % i for i in range(n)]
    return d, q


@pytest.mark.parametrize("problem,picks", [("roberts", "default+ncfn"), ("lorenz", "spread")])
def test_every_attempt_of_a_traced_instance_is_bit_identical(problem, picks):
    """The strongest parity statement the engine can make: for traced instances inside a larger ensemble, the COMPLETE
    solver state at the top of EVERY step attempt (divided differences, coefficients, error / weight / state vectors,
    all scalars, the LU factors, pivots and every counter) equals the oracle's, bit for bit -- the reference author's
    data_trace workflow (scripts/data_trace.ipynb: 377 frames for the default Roberts run) as a GPU-vs-oracle diff."""
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200.problems import lorenz_ensemble_params, roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 300
    if problem == "roberts":
        cls, kind, rtol, atol, touts = r.Roberts, ob.PK_ROBERTS, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS
        p = roberts_ensemble_params(B, seed=1234)
        p[:, 7] = r.Roberts.DEFAULT  # the reference's own example instance
    else:
        cls, kind, rtol, atol, touts = r.Lorenz63, ob.PK_LORENZ63, LORENZ_RTOL, LORENZ_ATOL, LORENZ_TOUTS
        p = lorenz_ensemble_params(B, seed=1234)
    yy0, yp0 = cls.consistent_ic(p)
    ref = ob.ensemble_solve(kind, 3, p, yy0, yp0, rtol, atol, touts, nthreads=8, want_traj=False)
    picked = [7, 0, B - 1]
    if problem == "roberts":
        picked += [int(i) for i in np.nonzero(ref["counters"]["ncfn"] > 0)[0][:2]]  # Newton-recover instances
        picked += [int(np.argmax(ref["counters"]["netf"]))]
    dn, qn = _trace_names(3)
    for inst in picked:
        d_ref, q_ref, _ = ob.attempt_trace(kind, p[:, inst], yy0[:, inst], yp0[:, inst], rtol, atol, touts)
        ida = r.Ida(cls(p), yy0, yp0, r.TolControlSV(rtol, atol))
        tr = ida.trace(inst)
        ida.solve_ladder(touts, want_traj=False)
        d, q = tr.rows()
        ida.close()
        assert d.shape == d_ref.shape and q.shape == q_ref.shape, "attempt counts differ for instance %d" % inst
        assert d.shape[0] == ref["counters"]["nattempts"][inst]
        bad_q = np.argwhere(q != q_ref)
        assert bad_q.size == 0, "instance %d: int field %s differs at attempt %d" % (inst, qn[bad_q[0][1]], bad_q[0][0])
        bad_d = np.argwhere(d.view(np.uint64) != d_ref.view(np.uint64))
        assert bad_d.size == 0, "instance %d: field %s differs at attempt %d (%r vs %r)" % (
            inst, dn[bad_d[0][1]], bad_d[0][0], d[tuple(bad_d[0])], d_ref[tuple(bad_d[0])])
    if problem == "roberts":
        d7, _, _ = ob.attempt_trace(kind, p[:, 7], yy0[:, 7], yp0[:, 7], rtol, atol, touts)
        assert d7.shape[0] == 377  # the frame count of the reference author's notebook


@pytest.mark.parametrize("M,kernel", [(10, "cta"), (4, "warp"), (6, "cta")])
def test_every_attempt_of_a_traced_heat_instance_is_bit_identical(M, kernel):
    """the same per-attempt statement for the team kernels on the dense heat problem: at the top of every attempt the
    whole state INCLUDING the n x n LU factors and the pivots (blocked look-ahead LU for n = 100 / 36, group LU for
    n = 16) equals the oracle's bit for bit"""
    import rust_ida_b200 as r
    from rust_ida_b200.problems import heat_ensemble_params
    ob, _ = parity_oracle()
    n, B = M * M, 24
    p = heat_ensemble_params(B, M=M, seed=1234)
    yy0, yp0 = r.Heat2D.consistent_ic(p, M)
    touts = 0.01 * 2.0 ** np.arange(11)
    rtol, atol = 1e-6, np.full(n, 1e-8)
    dn, qn = _trace_names(n)
    for inst in (0, 13, B - 1):
        d_ref, q_ref, _ = ob.attempt_trace(ob.PK_HEAT2D, p[:, inst], yy0[:, inst], yp0[:, inst], rtol, atol, touts)
        ida = r.Ida(r.Heat2D(p, M=M), yy0, yp0, r.TolControlSV(rtol, atol))
        ida.set_kernel(kernel)
        tr = ida.trace(inst, max_attempts=1024)
        ida.solve_ladder(touts, want_traj=False)
        d, q = tr.rows()
        ida.close()
        assert d.shape == d_ref.shape and q.shape == q_ref.shape and d.shape[0] > 50
        bad_q = np.argwhere(q != q_ref)
        assert bad_q.size == 0, "instance %d: int field %s differs at attempt %d" % (inst, qn[bad_q[0][1]], bad_q[0][0])
        bad_d = np.argwhere(d.view(np.uint64) != d_ref.view(np.uint64))
        assert bad_d.size == 0, "instance %d: field %s differs at attempt %d (%r vs %r)" % (
            inst, dn[bad_d[0][1]], bad_d[0][0], d[tuple(bad_d[0])], d_ref[tuple(bad_d[0])])


# ---- rootfinding: src/impl_r_check.rs, the Root returns of Ida::solve, get_root_info / get_num_g_evals ------------------
def test_roberts_root_events_match_the_oracle():
    """GPU (ROOTS instantiation of the thread kernel, caller loop of examples/roberts.rs:117-140 on the device) vs the
    oracle's restatement of IDARcheck1/2/3 + IDARootfind, instance by instance: event times and get_root_info
    bit-identical, get_num_g_evals equal, and the integration itself untouched by rootfinding."""
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200.problems import roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 700
    p = roberts_ensemble_params(B, seed=1234)
    p[:, 11] = r.Roberts.DEFAULT
    yy0, yp0 = r.Roberts.consistent_ic(p)
    plain = run_engine(r.Roberts, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, ladder=True)
    ida = r.Ida(r.Roberts(p), yy0, yp0, r.TolControlSV(ROBERTS_RTOL, ROBERTS_ATOL))
    ev = ida.enable_roots(max_events=8)
    status, traj = ida.solve_ladder(ROBERTS_TOUTS)
    e = ev.numpy()
    yy = np.asarray(ida.get_yy())
    cnt = {k: np.asarray(v) for k, v in ida.counters().items()}
    ida.close()
    # rootfinding never alters the step sequence or the returned solution at the output times
    assert (np.asarray(status) == 0).all()
    assert_bit_equal(yy, plain["yy"], "final state with / without rootfinding")
    assert_bit_equal(np.asarray(traj), plain["traj"], "y(tout) with / without rootfinding")
    for k in COUNTERS:
        assert np.array_equal(cnt[k], plain["counters"][k]), k
    assert (e["count"] == 2).all()  # y3 rises through 0.01, later y1 falls through 1e-4
    for inst in list(range(0, B, 7)) + [11, B - 1]:
        ref = ob.roots_run(ob.PK_ROBERTS, p[:, inst], yy0[:, inst], yp0[:, inst], ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS)
        assert e["count"][inst] == ref["t"].size
        assert_bit_equal(e["t"][:ref["t"].size, inst], ref["t"], "root times of instance %d" % inst)
        assert_bit_equal(e["iroots"][:ref["t"].size, :, inst], ref["iroots"], "get_root_info of instance %d" % inst)
        assert e["nge"][inst] == ref["counters"]["nge"]
    # the reference example's own instance: the two roots of SUNDIALS' idaRoberts_dns (t = 2.6400e-01, 2.0786e+07), with the
    # reference's sign convention for get_root_info (signum of g at the left end, src/impl_r_check.rs:388,569)
    assert abs(e["t"][0, 11] - 2.6400e-01) < 5e-5 and abs(e["t"][1, 11] / 2.0786e+07 - 1.0) < 2e-4
    assert e["iroots"][0, :, 11].tolist() == [0.0, -1.0] and e["iroots"][1, :, 11].tolist() == [1.0, 0.0]


def test_root_state_survives_between_launches():
    """solve(tout) call by call (one launch per tout) must find the same events as the one-launch ladder: the root state
    (glo, gactive, tlo, irfnd, nge ...) is persistent device state like the rest of the solver"""
    import rust_ida_b200 as r
    from rust_ida_b200.problems import roberts_ensemble_params
    B = 96
    p = roberts_ensemble_params(B, seed=99)
    yy0, yp0 = r.Roberts.consistent_ic(p)
    out = []
    for split in (False, True):
        ida = r.Ida(r.Roberts(p), yy0, yp0, r.TolControlSV(ROBERTS_RTOL, ROBERTS_ATOL))
        ev = ida.enable_roots(max_events=8)
        if split:
            for t in ROBERTS_TOUTS:
                ida.solve_ladder(np.array([t]), want_traj=False)
        else:
            ida.solve_ladder(ROBERTS_TOUTS, want_traj=False)
        out.append(ev.numpy())
        ida.close()
    assert np.array_equal(out[0]["count"], out[1]["count"]) and np.array_equal(out[0]["nge"], out[1]["nge"])
    assert_bit_equal(out[0]["t"], out[1]["t"], "root times, ladder vs call by call")


@pytest.mark.parametrize("problem,device_buffers", [("roberts", True), ("roberts", False), ("heat", True)])
def test_get_dky_matches_the_oracle(problem, device_buffers):
    """Dense output: Ida::get_dky (src/lib.rs:424-529) evaluated by the device for every instance equals
    the oracle's bit for bit, for every derivative order, inside and outside each instance's last step."""
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200.problems import heat_ensemble_params, roberts_ensemble_params
    ob, _ = parity_oracle()
    if problem == "roberts":
        B, n = 777, 3
        p = roberts_ensemble_params(B, seed=99)
        yy0, yp0 = r.Roberts.consistent_ic(p)
        rtol, atol, touts, kind = ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS[:6], ob.PK_ROBERTS
        mk = lambda pp: r.Roberts(pp)
    else:
        M, B = 10, 37
        n = M * M
        p = heat_ensemble_params(B, M=M, seed=5)
        yy0, yp0 = r.Heat2D.consistent_ic(p, M)
        rtol, atol, touts, kind = 1e-6, np.full(n, 1e-8), 0.01 * 2.0 ** np.arange(6), ob.PK_HEAT2D
        mk = lambda pp: r.Heat2D(pp, M=M)
    tend = float(touts[-1])
    # query times: the output time itself (inside every instance's last step), slightly before, far before (BadTimeValue
    # for most), beyond t_n (extrapolation is legal); orders 0..5 (BadK where k > kused of that instance)
    tq = np.array([tend] * 6 + [tend * (1 - 1e-3)] * 3 + [tend * 0.5, tend * 1.01, tend * 1.01])
    kq = np.array([0, 1, 2, 3, 4, 5, 0, 1, 2, 0, 0, 1], dtype=np.int32)
    ref, rst = ob.ensemble_dky(kind, n, p, yy0, yp0, rtol, atol, touts, tq, kq, nthreads=8)
    if device_buffers:
        ida = r.Ida(mk(torch.from_numpy(p).cuda()), torch.from_numpy(yy0).cuda(), torch.from_numpy(yp0).cuda(),
                    r.TolControlSV(rtol, atol))
    else:
        ida = r.Ida(mk(p), yy0, yp0, r.TolControlSV(rtol, atol))
    status, _ = ida.solve_ladder(touts, want_traj=False)
    seen = set()
    for q in range(len(tq)):
        dky, st = ida.get_dky(tq[q], int(kq[q]))
        if device_buffers:
            dky, st = dky.cpu().numpy(), st.cpu().numpy()
        assert np.array_equal(st, rst[q]), (q, np.flatnonzero(st != rst[q])[:5])
        ok = st == 0
        assert_bit_equal(dky[:, ok], ref[q][:, ok], "dky of query %d" % q)
        assert np.isnan(dky[:, ~ok]).all()  # untouched
        seen |= set(np.unique(st).tolist())
    # (the heat instances all end at order 5, so no k is too large for them -- and k = 0 runs the clamped recurrence)
    assert seen == ({0, -8, -10} if problem == "roberts" else {0, -8})
    # k = 0 at the output time is the solution the solver returned there, to rounding (a different recurrence)
    y_at, _ = ida.get_dky(tend, 0)
    yy = ida.get_yy()
    if device_buffers:
        y_at, yy = y_at.cpu().numpy(), yy.cpu().numpy()
    np.testing.assert_allclose(y_at, yy, rtol=1e-9, atol=1e-14)
    ida.close()


# ---- ida_id / ida_suppressalg inside the stepper -------------------------------------------------
@pytest.mark.parametrize("kernel", ["thread", "warp"])
def test_suppressalg_roberts_matches_the_oracle(kernel):
    """Roberts' third equation is algebraic (y1 + y2 + y3 = 1): with id = [1, 1, 0] and suppressalg the local-error
    norms (src/lib.rs:1353-1370 -> norm_wrms_masked, src/norm_rms.rs:49-57) leave y3 out.  Same bits as the oracle, and
    a different integration from the unmasked one (the option really reaches the step-size controller)."""
    from rust_ida_b200 import Roberts
    from rust_ida_b200.problems import roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 777 if kernel == "thread" else 96
    p = roberts_ensemble_params(B, seed=77)
    yy0, yp0 = Roberts.consistent_ic(p)
    ob.set_id([1, 1, 0], True)
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, nthreads=8, want_traj=True)
    ob.set_id(None)
    plain = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, nthreads=8)
    g = run_engine(Roberts, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, ladder=True, device_buffers=True,
                   kernel=kernel, id_mask=[1, 1, 0])
    compare(g, ref)
    assert (g["status"] == 0).all()
    assert (ref["counters"]["nst"] != plain["counters"]["nst"]).mean() > 0.5


def test_suppressalg_heat_boundary_rows_match_the_oracle():
    """the heat problem's boundary rows are algebraic (u = 0): masked out of the error norms, CTA-per-system kernel.
    Their error components are exactly zero, so masking them alone cannot change a norm; one interior grid row is masked
    as well to make the option observable in the step counts."""
    from rust_ida_b200 import Heat2D
    from rust_ida_b200.problems import heat_ensemble_params
    ob, _ = parity_oracle()
    M, B = 10, 48
    n = M * M
    idm = np.zeros((M, M), dtype=np.uint8)
    idm[1:-1, 1:-1] = 1
    idm[5, :] = 0
    p = heat_ensemble_params(B, M=M, seed=3)
    yy0, yp0 = Heat2D.consistent_ic(p, M)
    touts = 0.01 * 2.0 ** np.arange(11)
    rtol, atol = 1e-6, np.full(n, 1e-8)
    ob.set_id(idm.ravel(), True)
    ref = ob.ensemble_solve(ob.PK_HEAT2D, n, p, yy0, yp0, rtol, atol, touts, nthreads=8, want_traj=True)
    ob.set_id(None)
    plain = ob.ensemble_solve(ob.PK_HEAT2D, n, p, yy0, yp0, rtol, atol, touts, nthreads=8)
    g = run_engine(Heat2D, p, yy0, yp0, rtol, atol, touts, ladder=True, kernel="cta", problem_kw={"M": M}, id_mask=idm.ravel())
    compare(g, ref)
    assert not np.array_equal(ref["counters"]["nst"], plain["counters"]["nst"])


def test_suppressalg_cannot_be_combined_with_roots():
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200._lib import B200IdaError
    p = np.array(r.Roberts.DEFAULT).reshape(3, 1).repeat(8, axis=1).copy()
    yy0, yp0 = r.Roberts.consistent_ic(p)
    ida = r.Ida(r.Roberts(torch.from_numpy(p).cuda()), torch.from_numpy(yy0).cuda(), torch.from_numpy(yp0).cuda(),
                r.TolControlSV(ROBERTS_RTOL, ROBERTS_ATOL))
    ida.enable_roots(8)
    ida.set_id([1, 1, 0], True)
    with pytest.raises(B200IdaError):
        ida.solve_ladder(ROBERTS_TOUTS)
    ida.close()


# ---- row f4: the reference's data_trace JSON from a device trace --------------------------------------------------------------
def test_device_trace_serialises_to_the_reference_json(tmp_path):
    import json
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200 import trace_json
    p = r.problems.roberts_ensemble_params(64, seed=1234)
    p[:, 5] = r.Roberts.DEFAULT
    yy0, yp0 = r.Roberts.consistent_ic(p)
    ida = r.Ida(r.Roberts(torch.from_numpy(p).cuda()), torch.from_numpy(yy0).cuda(), torch.from_numpy(yp0).cuda(),
                r.TolControlSV(ROBERTS_RTOL, ROBERTS_ATOL))
    tr = ida.trace(5)
    ida.solve_ladder(ROBERTS_TOUTS, want_traj=False)
    d, q = tr.rows()
    path = tmp_path / "roberts_b200.json"
    assert trace_json.write(str(path), d, q, 3, ROBERTS_RTOL, ROBERTS_ATOL) == 377
    data = json.load(open(path))["data"]
    nst = [s["counters"]["ida_nst"] for s in data]
    assert nst[0] == 0 and nst[-1] == 361 and data[-1]["nlp"]["lp"]["nje"] in (59, 60)  # frames precede their attempt
    ida.close()


# ---- re-grouping between segments of a solve: pure engine tuning, results must not move -----------------------
@pytest.mark.parametrize("trips,segments", [(16, 8), (128, 4), (0, 0)])
def test_regrouped_solve_is_bit_identical(trips, segments):
    """70,000 Roberts instances (enough for the segmented path), every output time's y and every counter against the
    oracle, for an aggressive re-grouping schedule, the default one, and re-grouping switched off; then a second ladder
    on the re-grouped slots and a re-initialisation (which must restore the parameter order) give the same again."""
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200.problems import roberts_ensemble_params
    ob, _ = parity_oracle()
    B = 70000
    p = roberts_ensemble_params(B, seed=4321)
    yy0, yp0 = r.Roberts.consistent_ic(p)
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS, nthreads=16, want_traj=True)
    dp, dy, dyp = (torch.from_numpy(v).cuda() for v in (p, yy0, yp0))
    ida = r.Ida(r.Roberts(dp), dy, dyp, r.TolControlSV(ROBERTS_RTOL, ROBERTS_ATOL))
    ida.set_regroup(trips, segments)
    for rep in range(2):
        if rep:
            ida.reinit(dy, dyp)
        # first half of the ladder, then the rest in a second call on whatever slot order the first one left
        st1, tr1 = ida.solve_ladder(ROBERTS_TOUTS[:5])
        st2, tr2 = ida.solve_ladder(ROBERTS_TOUTS[5:])
        traj = torch.cat([tr1, tr2]).cpu().numpy()
        assert (st1 == 0).all() and (st2 == 0).all()
        assert_bit_equal(traj, ref["yy_traj"], "yy_traj")
        assert_bit_equal(ida.get_yy().cpu().numpy(), ref["yy"], "yy")
        assert_bit_equal(ida.get_yp().cpu().numpy(), ref["yp"], "yp")
        c = ida.counters()
        for k in COUNTERS:
            assert np.array_equal(c[k].cpu().numpy(), ref["counters"][k]), k
        assert ida.fallback_count() == 0
    ida.close()
