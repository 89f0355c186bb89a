"""GPU parity tests for the sparse path (supernodal Cholesky of A diag(theta) A'), through the C ABI,
against the CPU oracle.  Mirrors sparse-newton-solve.lisp:177-269 (random 5% matrices with a forced
diagonal, inf-norm residuals <= 1e-6, leak check) and adds factor / counter / PDAS parity."""
import numpy as np
import pytest
import scipy.sparse as sp

from cholesky_is_magic_b200 import lpgen, nes, newton_solve, pdas
from cholesky_is_magic_b200.standard_form import StandardForm, Triplets
from oracle import newton_solve as ons
from oracle import pdas as opdas

pytestmark = pytest.mark.gpu


def to_device(common, A):
    A = sp.csc_matrix(A)
    A.sort_indices()
    return nes.Matrix.from_csc(common, A.indptr, A.indices, A.data, A.shape[0], A.shape[1])


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def banded(rng, m, n, bw, per_col):
    rows, cols, vals = [], [], []
    for j in range(n):
        c = j * m // n if j >= m else j
        r = np.unique(np.clip(c + rng.integers(-bw, bw + 1, per_col - 1), 0, m - 1))
        r = np.union1d(r, [c])
        rows += r.tolist(); cols += [j] * len(r); vals += (1 + rng.random(len(r))).tolist()
    return sp.csc_matrix((vals, (rows, cols)), shape=(m, n))


CASES = [("random", 1, 1), ("random", 7, 12), ("random", 30, 30), ("random", 60, 150), ("random", 300, 700),
         ("banded", 500, 1200), ("banded", 2000, 5000)]


def make(kind, m, n, seed=0):
    rng = np.random.default_rng(seed + m)
    if kind == "random":
        return ons.random_sparse_matrix(rng, m, n, 0.05 if m <= 60 else 0.01), rng
    return banded(rng, m, n, 40, 6), rng


@pytest.mark.parametrize("kind,m,n", CASES)
def test_sparse_factor_residual_counters_and_solve(common, kind, m, n):
    A, rng = make(kind, m, n)
    s = np.sqrt(0.1 + 10 * rng.random(n))
    b = rng.random(m)
    Ad = to_device(common, A)
    Ad.scale(s)
    L = nes.Factor(common, Ad)
    M = ons.normal_matrix(A, s)
    assert common.anz == np.count_nonzero(np.tril((abs(A) @ abs(A).T).toarray()))   # nnz(tril(A A'))
    assert common.aatfl == float((np.diff(A.indptr).astype(float) ** 2).sum())
    assert common.lnz >= common.anz and common.fl >= common.lnz
    assert L.factorize(Ad)
    # P M P' = L L'
    out = np.zeros((m, m), order="F")
    perm = np.zeros(m, dtype=np.int32)
    common.check(common.lib.nes_factor_to_dense(L.ptr, out.ctypes.data_as(nes._dp), m,
                                                perm.ctypes.data_as(nes._ip), common.ptr), "to_dense")
    assert sorted(perm.tolist()) == list(range(m))
    assert relerr(out @ out.T, M[np.ix_(perm, perm)]) <= 1e-12
    x = L.solve(b)
    assert np.linalg.norm(M @ x - b) / np.linalg.norm(b) <= 1e-10
    assert relerr(x, ons.solve_spd(M, b)) <= 1e-8
    L.free()
    Ad.free()


@pytest.mark.parametrize("m,n", [(5, 9), (120, 260)])
def test_sparse_sdmult(common, m, n):
    rng = np.random.default_rng(m)
    A = sp.random(m, n, density=0.1, random_state=3, format="csc")
    Ad = to_device(common, A)
    x, y = rng.standard_normal(n), rng.standard_normal(m)
    np.testing.assert_allclose(Ad.sdmult(x), A @ x, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Ad.sdmult(y, transpose=True), A.T @ y, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Ad.sdmult(x, y=y, alpha=-1.0, beta=0.5), 0.5 * y - A @ x, rtol=1e-12, atol=1e-13)
    s = 0.5 + rng.random(n)
    Ad.scale(s)
    np.testing.assert_allclose(Ad.sdmult(y, transpose=True), s * (A.T @ y), rtol=1e-12, atol=1e-13)
    Ad.free()


def test_triplets_sum_duplicates_like_cholmod(common):
    rows = [0, 1, 1, 0, 2]; cols = [0, 0, 0, 1, 2]; vals = [1.0, 2.0, 3.0, 4.0, 5.0]
    Ad = nes.Matrix.from_triplets(common, rows, cols, vals, 3, 3)
    assert Ad.nnz == 4
    np.testing.assert_allclose(Ad.sdmult(np.array([1.0, 0, 0])), [1.0, 5.0, 0.0])
    Ad.free()


def test_sparse_kkt_property_like_reference_test(common):
    """(test max) of sparse-newton-solve.lisp:260-269 incl. its leak check (the `common` fixture)."""
    rng = np.random.default_rng(77)
    worst = 0.0
    for m in list(range(1, 12)) + [25, 90]:
        for n in (m, m + 2, 2 * m + 1):
            l, u, w, z, A, e, f, g, h = ons.random_sparse_case(rng, m, n)
            Ad = to_device(common, A)
            res = newton_solve.test_kkt_solve(l, u, w, z, Ad, e, f, g, h, A, ord=np.inf)
            Ad.free()
            worst = max(worst, max(res))
    assert worst <= 1e-6


def test_sparse_kkt_matches_oracle_with_filters(common):
    rng = np.random.default_rng(5)
    l, u, w, z, A, e, f, g, h = ons.random_sparse_case(rng, 80, 200)
    u[::4] = 1e8
    Ad = to_device(common, A)
    got = newton_solve.solve_kkt_newton(l, u, w, z, Ad, e, f, g, h)        # filters default on for sparse
    want = ons.solve_kkt_newton(l, u, w, z, A, e, f, g, h, filters=True)
    Ad.free()
    for a, b in zip(got, want):
        assert relerr(a, b) <= 1e-9


def test_sparse_not_positive_definite(common):
    A = sp.csc_matrix(np.array([[1.0, 0, 0], [1.0, 0, 0], [0, 0, 2.0]]))   # rows 0 and 1 identical
    Ad = to_device(common, A)
    L = nes.Factor(common, Ad)
    assert not L.factorize(Ad)
    assert common.status == nes.NES_NOT_POSDEF
    L.free()
    Ad.free()


@pytest.mark.parametrize("m,n", [(60, 150), (400, 1000)])
def test_sparse_pdas_matches_oracle(common, m, n):
    sf = lpgen.sparse_lp(m, n, nnz_per_col=6, bandwidth=24, seed=1)
    A = sp.csc_matrix((sf.A.value, (sf.A.row, sf.A.col)), shape=(m, n))
    ost = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), A, sf.b, sf.l, sf.u)   # filters on (sparse)
    oobj, ogap, oit = opdas.pdas(ost, 400)
    st = pdas.make_pdas(sf)
    obj, gap, it = pdas.pdas(st, 400, native_loop=True)
    assert it == oit
    assert abs(obj - oobj) <= 1e-8 * abs(oobj)


def test_config4_flow_mps_to_pdas(common, tmp_path):
    """BASELINE config 4 end to end at test size: generator -> MPS file -> read-mps -> to-standard-form
    -> make-pdas -> pdas with the sparse Newton solve on the GPU, against the oracle on the same file."""
    from cholesky_is_magic_b200 import read_mps
    m, n = 120, 300
    sf0 = lpgen.sparse_lp(m, n, nnz_per_col=5, bandwidth=16, seed=7)
    path = tmp_path / "config4.mps"
    read_mps.write_mps(path, sf0)
    sf = read_mps.to_standard_form(read_mps.read_mps_file(path))
    A = sp.csc_matrix((sf.A.value, (sf.A.row, sf.A.col)), shape=(sf.ncons, sf.nvars))
    ost = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), A, sf.b, sf.l, sf.u)
    oobj, ogap, oit = opdas.pdas(ost, 400)
    obj, gap, it = pdas.pdas(pdas.make_pdas(sf), 400)
    assert it == oit and abs(obj - oobj) <= 1e-8 * abs(oobj)


@pytest.mark.parametrize("name", ["dense_row", "components", "dense_block"])
def test_sparse_edge_patterns(common, name):
    """A fully dense row, a forest of disconnected components, chains wider than one supernode."""
    from test_symbolic import _edge_matrices
    A = _edge_matrices()[name]
    m, n = A.shape
    rng = np.random.default_rng(1)
    s = np.sqrt(0.1 + 10 * rng.random(n))
    b = rng.random(m)
    common.lib.nes_set_ordering_leaf(common.ptr, 64)
    try:
        Ad = to_device(common, A)
        Ad.scale(s)
        L = nes.Factor(common, Ad)
        assert L.factorize(Ad)
        x = L.solve(b)
        M = ons.normal_matrix(A, s)
        assert np.linalg.norm(M @ x - b) <= 1e-12 * np.linalg.norm(M) * np.linalg.norm(x)
        out = np.zeros((m, m), order="F")
        perm = np.zeros(m, dtype=np.int32)
        common.check(common.lib.nes_factor_to_dense(L.ptr, out.ctypes.data_as(nes._dp), m,
                                                    perm.ctypes.data_as(nes._ip), common.ptr), "to_dense")
        Mp = M[np.ix_(perm, perm)]
        assert relerr(out @ out.T, Mp) <= 1e-12
        # refactorize with another scale through the same analysis (solve-sparse-recycle), bitwise repeatable
        x1 = L.solve(b)
        assert L.factorize(Ad)
        np.testing.assert_array_equal(L.solve(b), x1)
        L.free()
        Ad.free()
    finally:
        common.lib.nes_set_ordering_leaf(common.ptr, 0)


def test_sparse_empty_row_is_not_positive_definite(common):
    A = sp.csc_matrix(np.array([[1.0, 2.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 1.0]]))   # row 1 is empty
    Ad = to_device(common, A)
    L = nes.Factor(common, Ad)
    assert not L.factorize(Ad)
    assert common.status == nes.NES_NOT_POSDEF and 0 <= common.minor < 3
    L.free()
    Ad.free()
