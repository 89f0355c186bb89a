"""BASELINE.json full sizes, checked through size-independent properties (the oracle cannot run these
in seconds): normal-equation residuals computed through independent kernels (the GEMV path), KKT block
residuals of solve-kkt-newton (the reference's own test, newton-solve.lisp:166-182), determinism."""
import numpy as np
import pytest

from cholesky_is_magic_b200 import batched, lpgen, nes, newton_solve

pytestmark = pytest.mark.gpu


def test_config2_normal_equations_m8192_n16384(common):
    m, n = 8192, 16384
    A = nes.Matrix.generate_dense(common, m, n, 0)
    rng = np.random.default_rng(0)
    s = np.sqrt(0.1 + 10 * rng.random(n))
    b = rng.random(m)
    A.scale(s)
    L = nes.Factor(common, A)
    assert L.factorize(A)
    x = L.solve(b)
    # (A s)(A s)' x - b through the GEMV kernels, not through the factor
    r = A.sdmult(A.sdmult(x, transpose=True)) - b
    assert np.linalg.norm(r) / np.linalg.norm(b) <= 1e-10
    # bitwise reproducible: same inputs, same factor and solution
    assert L.factorize(A)
    x2 = L.solve(b)
    np.testing.assert_array_equal(x, x2)
    # spot-check the factor residual on a leading block: (L L')[:k,:k] == M[:k,:k]
    k = 300
    Lh = L.to_dense()[:k, :k]
    Ah = lpgen.dense_matrix(k, n, 0)           # first k rows of the generated matrix
    M = (Ah * s ** 2) @ Ah.T
    assert np.linalg.norm(Lh @ Lh.T - M) / np.linalg.norm(M) <= 1e-12
    # and on the trailing block (last panels of the look-ahead schedule)
    assert common.form_flops == float(m) * m * n
    rows = np.arange(m - k, m, dtype=np.uint64)[:, None]
    At = lpgen.dense_entry(0, rows, np.arange(n, dtype=np.uint64)[None, :])
    At[np.arange(k), np.arange(m - k, m)] += 1.0
    Mt = (At * s ** 2) @ At.T
    Lt = L.to_dense()[m - k:, :]
    assert np.linalg.norm(Lt @ Lt.T - Mt) / np.linalg.norm(Mt) <= 1e-12
    L.free()
    A.free()


def test_config2_kkt_block_residuals(common):
    """test-kkt-solve at m=8192, n=16384 with the reference's data distributions."""
    m, n = 8192, 16384
    A = nes.Matrix.generate_dense(common, m, n, 1)
    rng = np.random.default_rng(1)
    l, u, w, z = (0.1 + 10 * rng.random(n) for _ in range(4))
    e, f, h = rng.random(n), rng.random(n), rng.random(n)
    g = rng.random(m)
    dw, dx, dy, dz = newton_solve.solve_kkt_newton(l, u, w, z, A, e, f, g, h)
    r1 = np.linalg.norm(u * dw - w * dx - e)
    r2 = np.linalg.norm(z * dx + l * dz - f)
    r3 = np.linalg.norm(A.sdmult(dx) - g)
    r4 = np.linalg.norm(A.sdmult(dy, transpose=True) + dz - dw - h)
    assert max(r1, r2, r3, r4) <= 1e-6          # the reference's report threshold
    A.free()


def test_config5_batch_1024_of_m256(common):
    B, m, n = 1024, 256, 512
    rng = np.random.default_rng(5)
    A = rng.random((B, m, n))
    A[:, np.arange(m), np.arange(m)] += 1.0
    s = np.sqrt(0.1 + 10 * rng.random((B, n)))
    rhs = rng.random((B, m))
    bt = batched.Batch(A)
    x, status = bt.normal_solve(s, rhs)
    bt.free()
    assert not status.any()
    for b in (0, 1, 511, 1023):
        M = (A[b] * s[b] ** 2) @ A[b].T
        assert np.linalg.norm(M @ x[b] - rhs[b]) / np.linalg.norm(rhs[b]) <= 1e-10


def test_config4_sparse_m100k_n250k(common):
    m, n = 100_000, 250_000
    sf = lpgen.sparse_lp(m, n, nnz_per_col=10, bandwidth=200, seed=0)
    A = nes.Matrix.from_triplets(common, sf.A.row, sf.A.col, sf.A.value, m, n)
    rng = np.random.default_rng(4)
    A.scale(np.sqrt(0.1 + 10 * rng.random(n)))
    L = nes.Factor(common, A)
    assert common.lnz >= common.anz > m
    assert L.factorize(A)
    b = rng.random(m)
    x = L.solve(b)
    r = A.sdmult(A.sdmult(x, transpose=True)) - b
    assert np.linalg.norm(r) / np.linalg.norm(b) <= 1e-8
    # bitwise reproducible (owner-computes extend-add, fixed-order reductions), analysis reused
    assert L.factorize(A)
    np.testing.assert_array_equal(L.solve(b), x)
    # linearity of the solve: (M^-1)(2 b + c) = 2 M^-1 b + M^-1 c up to rounding
    c2 = rng.random(m)
    lhs = L.solve(2 * b + c2)
    rhs = 2 * x + L.solve(c2)
    assert np.linalg.norm(lhs - rhs) <= 1e-9 * np.linalg.norm(rhs)
    # the counters the reference prints (affine-scaling.lisp:273-279) are consistent
    assert common.fl >= common.lnz >= common.anz and common.aatfl == float(
        (np.bincount(sf.A.col, minlength=n).astype(float) ** 2).sum())
    L.free()
    A.free()


def test_config4_scale_first_order_solver_properties(common):
    """APPROX at config-4 scale (850k stacked variables): the gradient is consistent with the value along a
    random direction, a run is bitwise repeatable, and the penalised objective goes down."""
    from cholesky_is_magic_b200 import approx as gap
    m, n = 100_000, 250_000
    sf = lpgen.sparse_lp(m, n, nnz_per_col=10, bandwidth=200, seed=0)
    st = gap.make_approx(sf)
    try:
        rng = np.random.default_rng(0)
        x = rng.standard_normal(st.nvars)
        d = rng.standard_normal(st.nvars)
        f0, g, _ = gap.value_and_gradient(st, x)
        h = 1e-6
        fp = gap.value_and_gradient(st, x + h * d)[0]
        fm = gap.value_and_gradient(st, x - h * d)[0]
        assert abs((fp - fm) / (2 * h) - g @ d) <= 1e-5 * abs(g @ d)
        v0 = gap.value_and_gradient(st, np.clip(np.zeros(st.nvars), st.l, st.u))[0]
        z1, it1, r1, s1 = gap.approx(st, 200)
        z2, it2, r2, s2 = gap.approx(st, 200)
        np.testing.assert_array_equal(z1, z2)
        assert it1 == it2 == 200 and s1[3] < 0.5 * v0
        assert np.all(z1 >= st.l) and np.all(z1 <= st.u)
    finally:
        st.free()
