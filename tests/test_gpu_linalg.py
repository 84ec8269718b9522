"""GPU parity: batched getrf / getrs / wrms through the C ABI vs the CPU oracle -- bit-exact factors,
pivots (LU pivot sequences must be bit-exact), solutions and norms."""
import json
import os

import numpy as np
import pytest

from helpers import assert_bit_equal

pytestmark = pytest.mark.gpu


def gen_matrices(n, batch, seed=1234):
    """the reference's own (commented-out) generator: U(0, 1/n) + anti-identity (benches/dense.rs:236-255)"""
    rng = np.random.default_rng(seed + n)
    a = rng.uniform(0.0, 1.0 / n, size=(batch, n, n))
    a += np.eye(n)[::-1][None]
    b = rng.uniform(0.0, 1.0, size=(batch, n))
    return a, b


def oracle_lu(a, b):
    from oracle import binding as ob
    lus, pivs, xs, infos = [], [], [], []
    for k in range(a.shape[0]):
        lu, piv, info = ob.getrf(a[k])
        lus.append(lu)
        pivs.append(piv)
        infos.append(info)
        xs.append(ob.getrs(lu, piv, b[k]) if info == 0 else b[k])
    return np.array(lus), np.array(pivs), np.array(xs), np.array(infos)


def run_gpu(a, b, layout, device_buffers, **dense_kw):
    """a: (batch, n, n) mathematical matrices, b: (batch, n)"""
    import torch
    import rust_ida_b200 as r
    batch, n, _ = a.shape
    if layout == "soa":
        A = np.ascontiguousarray(a.transpose(1, 2, 0))  # [i, j, b]
        X = np.ascontiguousarray(b.T)
    else:
        A = np.ascontiguousarray(a.transpose(0, 2, 1))  # column-major per matrix: [b, j, i]
        X = b.copy()
    Bv = X.copy()
    ls = r.Dense.new(n, batch, layout=layout, raise_on_singular=False, **dense_kw)
    if device_buffers:
        A, X, Bv = (torch.from_numpy(v).cuda() for v in (A, X, Bv))
    ls.setup(A)
    ls.solve(A, X, Bv, 0.0)
    if device_buffers:
        torch.cuda.synchronize()
        A, X = A.cpu().numpy(), X.cpu().numpy()
        piv, info = ls.pivots.cpu().numpy(), ls.info.cpu().numpy()
    else:
        piv, info = ls.pivots, ls.info
    if layout == "soa":
        return A.transpose(2, 0, 1), piv.T, X.T, info
    return A.transpose(0, 2, 1), piv, X, info


@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("n,batch", [(3, 4099), (4, 1000), (8, 517), (16, 130), (16, 131), (16, 1), (32, 67), (64, 19), (100, 9),
                                     (128, 5), (200, 3), (256, 2)])
def test_getrf_getrs_bit_exact(n, batch, layout):
    a, b = gen_matrices(n, batch)
    lu_ref, piv_ref, x_ref, info_ref = oracle_lu(a, b)
    lu, piv, x, info = run_gpu(a, b, layout, device_buffers=(n % 2 == 0))
    assert np.array_equal(info, info_ref)
    assert np.array_equal(piv, piv_ref), "pivot sequences differ"
    assert_bit_equal(lu, lu_ref, "LU factors n=%d" % n)
    assert_bit_equal(x, x_ref, "solutions n=%d" % n)
    assert (piv[:, 0] != 0).any()  # the anti-identity forces row swaps


def test_golden_matrices_on_gpu(golden_dir):
    """the reference's own known answers (crates/linear/src/dense.rs:267-311), bit-exact incl. pivots [2,1,2]"""
    with open(os.path.join(golden_dir, "dense.json")) as f:
        t = json.load(f)["tests"]
    mats, expects = [], []
    for name in ("test_get_rf1", "test_get_rf2"):
        it = {i["name"]: i for i in t[name]}
        mats.append(np.array(it["mat_a"]["value"]).T)
        expects.append(np.array(it["expect"]["value"]).T)
    a = np.array(mats)
    for layout in ("soa", "aos"):
        lu, piv, _, info = run_gpu(a, np.ones((2, 3)), layout, device_buffers=False)
        assert info.tolist() == [0, 0]
        assert piv.tolist() == [[2, 1, 2], [2, 1, 2]]
        assert_bit_equal(lu, np.array(expects), "golden LU")
    # getrs known answers (dense.rs:215-265)
    import rust_ida_b200 as r
    for name in ("test_get_rs1", "test_get_rs2"):
        it = {i["name"]: i for i in t[name]}
        lu = np.array(it["mat_a"]["value"]).T
        ls = r.Dense.new(3, 1, layout="soa")
        ls.pivots = np.array(it["pivot"]["value"], dtype=np.int64).reshape(3, 1).copy()
        x = np.zeros((3, 1))
        ls.solve(np.ascontiguousarray(lu.reshape(3, 3, 1)), x, np.array(it["b"]["value"]).reshape(3, 1).copy(), 0.0)
        assert_bit_equal(x[:, 0], np.array(it["expect"]["value"]), name)


@pytest.mark.parametrize("layout,n", [("soa", 3), ("aos", 3), ("soa", 20), ("aos", 40), ("soa", 70), ("aos", 130), ("aos", 200)])
def test_singular_reports_one_based_column(layout, n):
    import rust_ida_b200 as r
    rng = np.random.default_rng(7)
    a = rng.uniform(0.5, 1.5, size=(4, n, n)) + np.eye(n)[None] * n
    a[1, :, 0] = 0.0            # first column zero -> col 1
    a[2, :, n - 1] = a[2, :, 0]  # last column duplicates the first -> zero pivot at the last step (col n)
    a[3] = 0.0
    _, _, _, info_ref = oracle_lu(a, np.ones((4, n)))
    _, _, _, info = run_gpu(a, np.ones((4, n)), layout, device_buffers=False)
    assert info.tolist() == info_ref.tolist()
    assert info[0] == 0 and info[1] == 1 and info[3] == 1
    ls = r.Dense.new(n, 4, layout=layout)
    A = np.ascontiguousarray(a.transpose(1, 2, 0) if layout == "soa" else a.transpose(0, 2, 1))
    with pytest.raises(r.LUFactFail) as ei:
        ls.setup(A)
    assert ei.value.col == 1 and ei.value.index == 1  # 1-based column, like linear::Error::LUFactFail


@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("n,batch", [(3, 1000), (32, 257), (100, 33)])
def test_wrms_bit_exact(n, batch, layout):
    import rust_ida_b200 as r
    from oracle import binding as ob
    rng = np.random.default_rng(n)
    x = rng.normal(size=(batch, n)) * 10.0 ** rng.integers(-8, 8, size=(batch, 1))
    w = 1.0 / (1e-4 * np.abs(x) + 1e-8)
    ref = np.array([ob.wrms(x[k], w[k]) for k in range(batch)])
    if layout == "soa":
        out = r.norm_wrms(np.ascontiguousarray(x.T), np.ascontiguousarray(w.T), layout="soa")
    else:
        out = r.norm_wrms(x, w, layout="aos")
    assert_bit_equal(out, ref, "wrms")
    # the reference's own pin: wrms(-0.5, 0.5, n=32) == 0.25
    assert r.norm_wrms(np.full((32, 1), -0.5), np.full((32, 1), 0.5))[0] == 0.25


@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("n,batch", [(1, 1), (1, 77), (2, 1), (2, 33), (5, 1), (7, 129), (9, 3), (17, 2), (31, 5), (33, 1), (65, 2)])
def test_edge_sizes(n, batch, layout):
    """smallest systems, single-system batches and sizes that straddle the kernel dispatch boundaries"""
    a, b = gen_matrices(n, batch, seed=99)
    lu_ref, piv_ref, x_ref, info_ref = oracle_lu(a, b)
    lu, piv, x, info = run_gpu(a, b, layout, device_buffers=(batch % 2 == 1))
    assert np.array_equal(info, info_ref) and np.array_equal(piv, piv_ref)
    assert_bit_equal(lu, lu_ref, "LU n=%d" % n)
    assert_bit_equal(x, x_ref, "x n=%d" % n)


@pytest.mark.parametrize("layout,n", [("aos", 16), ("aos", 32), ("soa", 8), ("aos", 48), ("soa", 80), ("aos", 100), ("aos", 130),
                                      ("aos", 200)])
def test_pivot_ties_and_nan_follow_the_reference(layout, n):
    """ties pick the FIRST row with the largest |a_ik|; a NaN never wins the `>` test and a NaN in row k keeps l = k
    (crates/linear/src/dense.rs:109-116) -- checked on every kernel family"""
    rng = np.random.default_rng(3)
    a = rng.integers(-3, 4, size=(6, n, n)).astype(np.float64)  # many exact ties and zeros
    a += np.eye(n)[None] * 0.5
    a[1, 3, 0] = np.nan
    a[2, 0, 0] = np.nan
    a[3, n - 1, 2] = np.inf
    b = rng.normal(size=(6, n))
    lu_ref, piv_ref, x_ref, info_ref = oracle_lu(a, b)
    lu, piv, x, info = run_gpu(a, b, layout, device_buffers=False)
    assert np.array_equal(info, info_ref)
    ok = info_ref == 0
    assert np.array_equal(piv[ok], piv_ref[ok])
    assert_bit_equal(lu[ok], lu_ref[ok], "LU with ties/NaN")
    assert_bit_equal(x[ok], x_ref[ok], "x with ties/NaN")


@pytest.mark.parametrize("n", [40, 100, 130, 200])
def test_singular_leaves_the_reference_partial_factors(n):
    """after a zero pivot in column k the reference returns with steps < k applied (dense.rs:120-122); the blocked
    kernels (shared-memory and global-memory variants) must leave exactly that matrix and those pivots behind"""
    rng = np.random.default_rng(11)
    a = rng.uniform(0.5, 1.5, size=(3, n, n)) + np.eye(n)[None] * n
    a[0, :, n // 2 + 3] = 0.0             # a zero column stays zero (a_kj == 0 skips it): zero pivot inside a panel
    a[1, :, 8] = 0.0                      # zero pivot at the first column of the second panel
    a[2, :, n - 1] = a[2, :, n - 2]       # (nearly) dependent last columns: zero or tiny pivot at the last step
    b = np.ones((3, n))
    lu_ref, piv_ref, _, info_ref = oracle_lu(a, b)
    lu, piv, _, info = run_gpu(a, b, "aos", device_buffers=True)
    assert info.tolist() == info_ref.tolist() and info_ref[0] == n // 2 + 4 and info_ref[1] == 9
    assert_bit_equal(lu, lu_ref, "partial factors n=%d" % n)
    for k in range(3):
        m = info_ref[k] if info_ref[k] > 0 else n
        assert np.array_equal(piv[k, :m], piv_ref[k, :m])


@pytest.mark.parametrize("n", [64, 100, 256])
def test_getrs_unaligned_batch_takes_the_non_tma_path(n):
    """getrs_tma needs a 16-byte aligned batch (cp.async.bulk); a batch that starts 8 bytes into an allocation must
    still be solved, bit-exactly, by the register-streaming kernel"""
    import torch
    from rust_ida_b200 import _lib
    batch = 7
    a, b = gen_matrices(n, batch, seed=5)
    lu_ref, piv_ref, x_ref, _ = oracle_lu(a, b)
    L = _lib.lib()
    buf = torch.zeros(batch * n * n + 1, dtype=torch.float64, device="cuda")
    A = buf[1:]
    assert A.data_ptr() % 16 == 8
    A.copy_(torch.from_numpy(np.ascontiguousarray(lu_ref.transpose(0, 2, 1)).reshape(-1)))
    piv = torch.from_numpy(piv_ref.astype(np.int64).reshape(-1)).cuda()
    x = torch.from_numpy(b.copy().reshape(-1)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.b200ida_getrs_batched(A.data_ptr(), piv.data_ptr(), x.data_ptr(), n, batch, _lib.LAYOUT_AOS, _lib.MEM_DEVICE, st))
    torch.cuda.synchronize()
    assert_bit_equal(x.cpu().numpy().reshape(batch, n), x_ref, "unaligned getrs n=%d" % n)


@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("n,batch,dev", [(3, 1000, True), (32, 257, False), (100, 33, True)])
def test_wrms_masked_bit_exact(n, batch, layout, dev):
    """norm_wrms_masked (src/norm_rms.rs:49-57) incl. the reference's own pin (:73-86)"""
    import torch
    import rust_ida_b200 as r
    from oracle import binding as ob
    rng = np.random.default_rng(n + 1)
    x = rng.normal(size=(batch, n)) * 10.0 ** rng.integers(-8, 8, size=(batch, 1))
    w = 1.0 / (1e-4 * np.abs(x) + 1e-8)
    mask = rng.random(n) < 0.7
    ref = np.array([ob.wrms_masked(x[k], w[k], mask) for k in range(batch)])
    X, W = (np.ascontiguousarray(x.T), np.ascontiguousarray(w.T)) if layout == "soa" else (x, w)
    if dev:
        out = r.norm_wrms_masked(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), mask, layout=layout).cpu().numpy()
    else:
        out = r.norm_wrms_masked(X, W, mask, layout=layout)
    assert_bit_equal(out, ref, "wrms_masked")
    m = np.ones(32, dtype=bool)
    m[-1] = False
    fac = np.sqrt(np.float64(31) / np.float64(32))
    assert r.norm_wrms_masked(np.full((32, 1), -0.5), np.full((32, 1), 0.5), m)[0] == fac * 0.5 * 0.5


# ---- SURVEY.md row N3 / BASELINE config 4: FP64 tensor-core (DMMA) trailing update, the TOLERANCE mode ---------------------
@pytest.mark.parametrize("n,batch", [(128, 7), (192, 4), (250, 3), (256, 5)])
def test_getrf_tensor_core_update_within_tolerance(n, batch):
    """B200IDA_LU_DMMA: pivots equal the oracle's, factors within rel 1e-12 (config 4's bar for the DMMA path), P A = L U,
    and the mode really differs from the bit-exact default somewhere (it is a different rounding, not a no-op)."""
    import torch
    import rust_ida_b200 as r
    a, b = gen_matrices(n, batch, seed=99)
    lu_ref, piv_ref, x_ref, info_ref = oracle_lu(a, b)
    A = torch.from_numpy(np.ascontiguousarray(a.transpose(0, 2, 1))).cuda()
    ls = r.Dense.new(n, batch, layout="aos", raise_on_singular=False, tensor_core_update=True)
    ls.setup(A)
    X = torch.from_numpy(b.copy()).cuda()
    ls.solve(A, X, torch.from_numpy(b.copy()).cuda(), 0.0)
    lu = A.cpu().numpy().transpose(0, 2, 1)
    piv, info = ls.pivots.cpu().numpy(), ls.info.cpu().numpy()
    assert np.array_equal(info, info_ref) and (info == 0).all()
    assert np.array_equal(piv, piv_ref), "pivot sequences differ"
    scale = np.abs(lu_ref).max(axis=(1, 2), keepdims=True)
    assert (np.abs(lu - lu_ref) / scale).max() < 1e-12
    np.testing.assert_allclose(X.cpu().numpy(), x_ref, rtol=1e-9, atol=1e-12)
    for k in range(batch):  # P A = L U
        Lm = np.tril(lu[k], -1) + np.eye(n)
        Um = np.triu(lu[k])
        pa = a[k].copy()
        for i in range(n):
            pa[[i, piv[k, i]]] = pa[[piv[k, i], i]]
        assert np.abs(Lm @ Um - pa).max() < 1e-12
    assert not np.array_equal(lu.view(np.uint64), lu_ref.view(np.uint64))


def test_tensor_core_flag_is_ignored_where_it_does_not_apply():
    """odd n, small n and SoA keep the bit-exact path even with the flag set"""
    import rust_ida_b200 as r
    for n, layout in ((100, "aos"), (129, "aos"), (8, "soa")):
        a, b = gen_matrices(n, 3, seed=5)
        lu_ref, piv_ref, _, _ = oracle_lu(a, b)
        A = np.ascontiguousarray(a.transpose(1, 2, 0)) if layout == "soa" else np.ascontiguousarray(a.transpose(0, 2, 1))
        ls = r.Dense.new(n, 3, layout=layout, raise_on_singular=False, tensor_core_update=True)
        ls.setup(A)
        lu = A.transpose(2, 0, 1) if layout == "soa" else A.transpose(0, 2, 1)
        assert_bit_equal(lu, lu_ref, "n=%d %s" % (n, layout))


# ---- B200IDA_LU_TEAM_PANEL: the CTA-team IDA kernel's factorisation (panel shared by one warp per 32-row slot, work-queue
#      update; lu_blocked.cuh lub_factor / lub_panel_mw) driven through the Level-1 entry point, so that its tie / NaN / inf
#      exchange paths, its zero-pivot exits and every panel-warp count (1..4 warps, one or two row slots per lane) are
#      checked against the oracle and not only on the well-behaved heat Jacobians -------------------------------------------
@pytest.mark.parametrize("n,batch", [(33, 5), (40, 7), (64, 9), (65, 4), (96, 5), (100, 9), (128, 3), (129, 3), (160, 2)])
def test_team_panel_flavour_is_bit_exact(n, batch):
    a, b = gen_matrices(n, batch)
    lu_ref, piv_ref, x_ref, info_ref = oracle_lu(a, b)
    lu, piv, x, info = run_gpu(a, b, "aos", device_buffers=True, team_panel=True)
    assert (info == 0).all() and (info_ref == 0).all()
    assert np.array_equal(piv, piv_ref)
    assert_bit_equal(lu, lu_ref, "LU (team panel) n=%d" % n)
    assert_bit_equal(x, x_ref, "x (team panel) n=%d" % n)


@pytest.mark.parametrize("n", [48, 64, 100, 130, 160])
def test_team_panel_ties_nan_inf_follow_the_reference(n):
    rng = np.random.default_rng(3)
    a = rng.integers(-3, 4, size=(8, n, n)).astype(np.float64)  # many exact ties and zeros, across the panel's warps
    a += np.eye(n)[None] * 0.5
    a[1, 3, 0] = np.nan
    a[2, 0, 0] = np.nan
    a[3, n - 1, 2] = np.inf
    a[4, n // 2, 1] = np.nan       # NaN in another warp's row slot
    a[5, 40, 9] = -np.inf          # inf in the second panel
    a[6, :, 5] = np.abs(a[6, :, 5]).max() + 1.0  # a whole column of equal maxima: the tie spans every warp of the panel
    b = rng.normal(size=(8, n))
    lu_ref, piv_ref, x_ref, info_ref = oracle_lu(a, b)
    lu, piv, x, info = run_gpu(a, b, "aos", device_buffers=True, team_panel=True)
    assert np.array_equal(info, info_ref)
    ok = info_ref == 0
    assert np.array_equal(piv[ok], piv_ref[ok])
    assert_bit_equal(lu[ok], lu_ref[ok], "LU with ties/NaN (team panel)")
    assert_bit_equal(x[ok], x_ref[ok], "x with ties/NaN (team panel)")


@pytest.mark.parametrize("n", [40, 100, 130])
def test_team_panel_singular_leaves_the_reference_partial_factors(n):
    rng = np.random.default_rng(11)
    a = rng.uniform(0.5, 1.5, size=(3, n, n)) + np.eye(n)[None] * n
    a[0, :, n // 2 + 3] = 0.0
    a[1, :, 8] = 0.0
    a[2, :, n - 1] = a[2, :, n - 2]
    b = np.ones((3, n))
    lu_ref, piv_ref, _, info_ref = oracle_lu(a, b)
    lu, piv, _, info = run_gpu(a, b, "aos", device_buffers=True, team_panel=True)
    assert info.tolist() == info_ref.tolist() and info_ref[0] == n // 2 + 4 and info_ref[1] == 9
    assert_bit_equal(lu, lu_ref, "partial factors (team panel) n=%d" % n)
    for k in range(3):
        m = info_ref[k] if info_ref[k] > 0 else n
        assert np.array_equal(piv[k, :m], piv_ref[k, :m])
