"""Pins the CPU oracle against every golden vector the reference's own tests hold.

Fixtures: tests/golden/*.json, extracted from /root/reference by tools/extract_golden.py.
`assert_eq!` pins in the reference are checked bit-exact here; `assert_nearly_eq!` pins (values
printed with %.16e from a SUNDIALS C run) are checked to rel 1e-15 and, where the printed value
round-trips, bit-exact as well.
"""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob

INT_FIELDS = ("kk", "kused", "knew", "phase", "ns", "nst", "maxord")


def load(golden_dir, name):
    with open(os.path.join(golden_dir, name + ".json")) as f:
        return json.load(f)["tests"]


def norm_name(name):
    """`ida.nlp.lp.ida_cj` / `ida_cj` / `cj` -> `cj`"""
    name = name.split(".")[-1]
    return name[4:] if name.startswith("ida_") else name


def split_before_after(items, skip=()):
    """The reference tests list each state variable twice: pre-conditions, then expectations."""
    before, after = {}, {}
    for it in items:
        k = norm_name(it["name"])
        if k in skip:
            continue
        (after if k in before else before)[k] = it["value"]
    return before, after


def apply_state(u, state):
    for k, v in state.items():
        if k in INT_FIELDS:
            u.set(k, int(v))
        elif k in ("suppressalg",):
            continue
        else:
            u.set(k, v)


# ---- LU known answers: crates/linear/src/dense.rs:215-328 (assert_eq! => bit-exact) -------------------
@pytest.mark.parametrize("name", ["test_get_rf1", "test_get_rf2"])
def test_getrf_golden(golden_dir, name):
    it = {i["name"]: i for i in load(golden_dir, "dense")[name]}
    # literals are column lists then `.transpose()`d: literal rows are columns of the matrix
    a = np.array(it["mat_a"]["value"]).T
    expect = np.array(it["expect"]["value"]).T
    lu, piv, info = ob.getrf(a)
    assert info == 0
    assert piv.tolist() == [2, 1, 2]
    assert np.array_equal(lu, expect)  # bit-exact


@pytest.mark.parametrize("name", ["test_get_rs1", "test_get_rs2"])
def test_getrs_golden(golden_dir, name):
    it = {i["name"]: i for i in load(golden_dir, "dense")[name]}
    a = np.array(it["mat_a"]["value"]).T
    x = ob.getrs(a, np.array(it["pivot"]["value"], dtype=np.int64), it["b"]["value"])
    assert np.array_equal(x, np.array(it["expect"]["value"]))  # bit-exact


def test_dense_4x4(golden_dir):
    it = {i["name"]: i for i in load(golden_dir, "dense")["test_dense1"]}
    a = np.array(it["mat_a"]["value"])  # not transposed in the reference
    lu, piv, info = ob.getrf(a)
    assert info == 0
    x = ob.getrs(lu, piv, it["b"]["value"])
    np.testing.assert_allclose(x, it["expected"]["value"], rtol=1e-9)


def test_getrf_singular_reports_one_based_column():
    a = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [1.0, 0.0, 1.0]])  # rank 2: zero pivot at k=2
    _, _, info = ob.getrf(a)
    assert info == 3  # dense.rs:120-122 Err(k+1)
    _, _, info = ob.getrf(np.zeros((3, 3)))
    assert info == 1


# ---- norms: src/norm_rms.rs:65-70, crates/nonlinear/src/norm_wrms.rs:41-50 ---------------------------
def test_wrms_golden():
    assert ob.wrms(np.full(32, -0.5), np.full(32, 0.5)) == 0.25


def test_wrms_masked_golden():
    """src/norm_rms.rs:73-86: all components but the last, expect fac * 0.5 * 0.5 with assert_eq!"""
    n = 32
    fac = np.sqrt(np.float64(n - 1) / np.float64(n))
    mask = np.ones(n, dtype=bool)
    mask[-1] = False
    assert ob.wrms_masked(np.full(n, -0.5), np.full(n, 0.5), mask) == fac * 0.5 * 0.5
    assert ob.wrms_masked(np.full(n, -0.5), np.full(n, 0.5), np.ones(n, dtype=bool)) == 0.25


# ---- Newton known answer: crates/nonlinear/src/newton.rs:306-343 -------------------------------------
def test_newton_golden(golden_dir):
    it = {i["name"]: i["value"] for i in load(golden_dir, "newton")["test_newton"]}
    r, y, niters, nconvfails = ob.newton_quad3(it["y0"], it["w"], 1e-2, 10)
    assert r == 0
    np.testing.assert_allclose(y - np.array(it["y_exp"]), it["expected_err"], rtol=1e-5)
    assert nconvfails == 0 and 1 <= niters <= 10


# ---- complete_step x3: src/tests/complete_step.rs (assert_eq! => bit-exact, rr through pow) ----------
@pytest.mark.parametrize("name", ["test1", "test2", "test3"])
def test_complete_step_golden(golden_dir, name):
    items = load(golden_dir, "complete_step")[name]
    before, after = split_before_after(items)
    err_k, err_km1 = before.pop("err_k"), before.pop("err_km1")
    u = ob.Unit(3)
    apply_state(u, before)
    u.complete_step(err_k, err_km1)
    for k, v in after.items():
        got = u.get(k)
        if k in INT_FIELDS:
            assert got == int(v), k
        else:
            assert np.array_equal(np.asarray(got), np.asarray(v)), (k, got, v)  # bit-exact


def test_complete_step_rr_is_pow_pinned(golden_dir):
    items = load(golden_dir, "complete_step")["test3"]
    _, after = split_before_after(items)
    assert after["rr"] == 2.2575213239991561  # src/tests/complete_step.rs:289


# ---- set_coeffs x2: src/tests/set_coeffs.rs (assert_nearly_eq!) ---------------------------------------
@pytest.mark.parametrize("name", ["test1", "test2"])
def test_set_coeffs_golden(golden_dir, name):
    items = load(golden_dir, "set_coeffs")[name]
    before, after = split_before_after(items)
    ck_expect = before.pop("ck_expect", None) or after.pop("ck_expect")
    u = ob.Unit(3)
    apply_state(u, before)
    ck = u.set_coeffs()
    np.testing.assert_allclose(ck, ck_expect, rtol=1e-15)
    for k, v in after.items():
        got = u.get(k)
        if k in INT_FIELDS:
            assert got == int(v), k
        else:
            np.testing.assert_allclose(np.asarray(got), np.asarray(v), rtol=1e-15, atol=0, err_msg=k)


# ---- predict: src/tests/predict.rs ------------------------------------------------------------------
def test_predict_golden(golden_dir):
    items = load(golden_dir, "predict")["test1"]
    before, after = split_before_after(items)
    u = ob.Unit(3)
    apply_state(u, before)
    u.predict()
    for k, v in after.items():
        np.testing.assert_allclose(np.asarray(u.get(k)), np.asarray(v), rtol=1e-15, atol=0, err_msg=k)


# ---- restore: src/tests/restore.rs -------------------------------------------------------------------
def test_restore_golden(golden_dir):
    items = {norm_name(i["name"]): i["value"] for i in load(golden_dir, "restore")["test_restore1"]}
    u = ob.Unit(3)
    for k in ("tn", "hh", "phi", "psi", "beta"):
        u.set(k, items[k])
    u.set("ns", int(items["ns"]))
    u.set("kk", int(items["kk"]))
    u.restore(items["saved_t"])
    assert u.get("tn") == items["saved_t"]
    assert u.get("ns") == 1 and u.get("kk") == 2
    np.testing.assert_allclose(u.get("beta"), items["beta_after"], rtol=1e-15)
    np.testing.assert_allclose(u.get("psi"), items["psi_after"], rtol=1e-15)
    np.testing.assert_allclose(u.get("phi"), items["phi_after"], rtol=1e-15)


# ---- test_error x2: src/tests/test_error.rs -----------------------------------------------------------
@pytest.mark.parametrize("name,expect_pass", [("test1", False), ("test2", True)])
def test_test_error_golden(golden_dir, name, expect_pass):
    it = {norm_name(i["name"]): i["value"] for i in load(golden_dir, "test_error")[name]}
    u = ob.Unit(3)
    u.set("kk", int(it["kk"]))
    for k in ("phi", "ee", "ewt", "sigma"):
        u.set(k, it[k])
    err_k, err_km1, conv = u.test_error(it["ck"])
    assert conv == expect_pass
    assert u.get("knew") == int(it["knew"])
    np.testing.assert_allclose(err_k, it["err_k"], rtol=1e-15)
    np.testing.assert_allclose(err_km1, it["err_km1"], rtol=1e-15)


# ---- get_solution: src/tests/get_solution.rs ----------------------------------------------------------
def test_get_solution_golden(golden_dir):
    it = {norm_name(i["name"]): i["value"] for i in load(golden_dir, "get_solution")["test_get_solution"]}
    u = ob.Unit(3)
    u.set("kused", int(it["kused"]))
    for k in ("hh", "tn", "hused", "phi", "psi"):
        u.set(k, it[k])
    assert u.get_solution(it["t"]) == 0
    np.testing.assert_allclose(u.get("yy"), it["yret_expect"], rtol=1e-15)
    np.testing.assert_allclose(u.get("yp"), it["ypret_expect"], rtol=1e-15)


def test_get_dky_properties(golden_dir):
    """get_dky (src/lib.rs:424-529) has no test or caller in the reference, so the restatement is pinned by what the
    function must satisfy on the state of the reference's own get_solution fixture: k = 0 at t_n is phi[0] exactly
    (every other coefficient is a product with delt + psi_{-1} = 0); k = 0 and k = 1 are the same polynomial and its
    derivative that get_solution evaluates (different recurrence, so equal to rounding, and get_solution IS pinned to
    the reference's expected values above); the k-th coefficient is k! / (psi_0 .. psi_{k-1}); BadK / BadTimeValue."""
    it = {norm_name(i["name"]): i["value"] for i in load(golden_dir, "get_solution")["test_get_solution"]}
    u = ob.Unit(3)
    kused = int(it["kused"])
    u.set("kused", kused)
    for k in ("hh", "tn", "hused", "phi", "psi"):
        u.set(k, it[k])
    phi, psi, tn = np.array(it["phi"], dtype=np.float64).reshape(6, 3), np.array(it["psi"], dtype=np.float64), float(it["tn"])
    r, d0 = u.get_dky(tn, 0)
    assert r == 0 and np.array_equal(d0, phi[0] + 0.0)
    t = float(it["t"])
    assert u.get_solution(t) == 0
    r0, d0 = u.get_dky(t, 0)
    r1, d1 = u.get_dky(t, 1)
    assert r0 == 0 and r1 == 0
    np.testing.assert_allclose(d0, u.get("yy"), rtol=1e-13)
    np.testing.assert_allclose(d1, u.get("yp"), rtol=1e-12)
    np.testing.assert_allclose(d0, it["yret_expect"], rtol=1e-13)
    np.testing.assert_allclose(d1, it["ypret_expect"], rtol=1e-12)
    # highest derivative: only c_k^(k) = k!/(psi_0...psi_{k-1}) survives, times phi[k]
    rk, dk = u.get_dky(t, kused)
    import math
    assert rk == 0
    np.testing.assert_allclose(dk, phi[kused] * (math.factorial(kused) / np.prod(psi[:kused])), rtol=1e-14)
    assert u.get_dky(t, kused + 1)[0] == -10 and u.get_dky(t, -1)[0] == -10       # BadK
    assert u.get_dky(tn - 10.0 * float(it["hused"]), 0)[0] == -8                    # BadTimeValue
    # kused = 5, k = 0: the reference indexes past its 6-entry coefficient array; the restatement stops at j = 5
    u.set("kused", 5)
    r5, d5 = u.get_dky(t, 0)
    assert r5 == 0 and np.isfinite(d5).all()


# ---- end to end: examples/roberts.rs (config 1) ------------------------------------------------------
ROBERTS_TABLE = [  # (nst, kused, hused) at tout = 0.4*10^i -- SURVEY.md section 6 / BASELINE.md section 2
    (29, 3, 8.80238e-02), (43, 4, 6.33772e-01), (68, 4, 3.19323e+00), (95, 4, 3.32007e+01),
    (126, 3, 3.14577e+02), (161, 5, 2.50577e+03), (202, 3, 2.63706e+04), (250, 3, 1.71865e+05),
    (293, 4, 2.36551e+06), (325, 4, 2.68085e+07), (348, 3, 7.43054e+08), (362, 2, 7.54805e+09)]


def test_roberts_end_to_end():
    touts = 0.4 * 10.0 ** np.arange(12)
    st, rows, c = ob.single_trace(ob.PK_ROBERTS, [0.04, 1e4, 3e7], [1, 0, 0], [-0.04, 0.04, 0], 1e-4,
                                  [1e-8, 1e-6, 1e-6], touts)
    assert st == 0
    # 377 attempts = frames printed by the reference author's trace notebook (scripts/data_trace.ipynb)
    assert c["nattempts"] == 377
    assert (c["nst"], c["nre"], c["nni"], c["nje"], c["netf"], c["ncfn"]) == (362, 537, 537, 60, 15, 0)
    for row, (nst, k, h) in zip(rows, ROBERTS_TABLE):
        assert int(row[0]) == nst and int(row[1]) == k
        np.testing.assert_allclose(row[2], h, rtol=1e-5)
    y = rows[-1, 3:]
    # bit-identical to the survey's independent line-by-line probe of the reference
    assert y.tolist() == [4.86408986935633547e-08, 1.94563604063205543e-13, 9.99999951358906736e-01]
    # check_ans: examples/roberts.rs:9-51
    ref = np.array([5.2083474251394888e-08, 2.0833390772616859e-13, 9.9999994791631752e-01])
    ewt = 1.0 / (1e-4 * np.abs(ref) + 10.0 * np.array([1e-8, 1e-6, 1e-6]))
    assert ob.wrms(y - ref, ewt) < 1.0


# ---- rootfinding (src/impl_r_check.rs): pinned by the example's own problem -------------------------------------------
def test_oracle_rootfinding_on_the_reference_example():
    """examples/roberts.rs with its root functions g0 = y1 - 1e-4, g1 = y3 - 0.01 (src/sample_problems/roberts.rs:99-113).
    The reference has no rootfinding test; the pins are (a) the two root locations SUNDIALS publishes for this very
    problem (idaRoberts_dns: t = 2.6400e-01 and 2.0786e+07), (b) g_i(y(t_root)) = 0 to the root tolerance, (c) the
    integration being untouched (362 steps / 377 attempts, same final state), (d) the reference's get_root_info sign
    convention, signum of g at the left end of the bracketing interval (src/impl_r_check.rs:388, 569)."""
    touts = 0.4 * 10.0 ** np.arange(12)
    k = [0.04, 1e4, 3e7]
    r = ob.roots_run(ob.PK_ROBERTS, k, [1, 0, 0], [-0.04, 0.04, 0], 1e-4, [1e-8, 1e-6, 1e-6], touts)
    assert (r["status"] == 0).all() and r["t"].size == 2
    assert abs(r["t"][0] - 2.6400e-01) < 5e-5 and abs(r["t"][1] / 2.0786e+07 - 1.0) < 2e-4
    assert abs(r["y"][0][2] - 0.01) < 1e-9 and abs(r["y"][1][0] - 1e-4) < 1e-11
    assert r["iroots"].tolist() == [[0.0, -1.0], [1.0, 0.0]]
    c = r["counters"]
    assert (c["nst"], c["nattempts"], c["nre"], c["nje"], c["netf"]) == (362, 377, 537, 60, 15)
    st, rows, _ = ob.single_trace(ob.PK_ROBERTS, k, [1, 0, 0], [-0.04, 0.04, 0], 1e-4, [1e-8, 1e-6, 1e-6], touts)
    assert np.array_equal(r["yy"].view(np.uint64), rows[-1][3:].view(np.uint64))  # same final state as without roots
    assert c["nge"] > c["nst"]  # one g evaluation per step at least (r_check3), plus the Illinois iterations
