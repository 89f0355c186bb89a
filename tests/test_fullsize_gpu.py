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
    L.free()
    A.free()
