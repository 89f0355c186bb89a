"""GPU parity for the batched path (BASELINE config 5): batched potrf/trsv and batched affine scaling
against the per-problem CPU oracle."""
import numpy as np
import pytest

from cholesky_is_magic_b200 import batched, lpgen
from oracle import affine_scaling as oa
from oracle import newton_solve as ons

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,m,n", [(1, 5, 9), (5, 40, 100), (3, 256, 512), (4, 200, 300), (2, 300, 400)])
def test_batched_normal_solve_matches_oracle(common, B, m, n):
    rng = np.random.default_rng(B * 1000 + m)
    A = rng.random((B, m, n)) + np.eye(m, n)[None]
    s = np.sqrt(0.1 + 10 * rng.random((B, n)))
    rhs = rng.random((B, m))
    bt = batched.Batch(A)
    x, status = bt.normal_solve(s, rhs)
    assert not status.any()
    for b in range(B):
        M = ons.normal_matrix(A[b], s[b])
        assert np.linalg.norm(M @ x[b] - rhs[b]) / np.linalg.norm(rhs[b]) <= 1e-10
        want = ons.solve_spd(M, rhs[b])
        assert np.linalg.norm(x[b] - want) / np.linalg.norm(want) <= 1e-8
    x2, _ = bt.normal_solve(None, rhs)       # unscaled
    for b in range(B):
        assert np.linalg.norm((A[b] @ A[b].T) @ x2[b] - rhs[b]) / np.linalg.norm(rhs[b]) <= 1e-10
    bt.free()


@pytest.mark.parametrize("B,m,n", [(1, 5, 9), (5, 40, 100), (3, 256, 512), (4, 200, 300), (2, 300, 400), (3, 64, 70)])
def test_batched_64_column_panels_opt_in_path(common, monkeypatch, B, m, n):
    """NES_BATCH_PANEL64: the 64 x 64 block potrf / trsm kernels (measured slower, kept opt-in): same gates,
    and the failing column of a singular problem is still reported per problem."""
    monkeypatch.setenv("NES_BATCH_PANEL64", "1")
    rng = np.random.default_rng(B * 77 + m)
    A = rng.random((B, m, n)) + np.eye(m, n)[None]
    s = np.sqrt(0.1 + 10 * rng.random((B, n)))
    rhs = rng.random((B, m))
    bt = batched.Batch(A)
    x, status = bt.normal_solve(s, rhs)
    assert not status.any()
    for b in range(B):
        M = ons.normal_matrix(A[b], s[b])
        assert np.linalg.norm(M @ x[b] - rhs[b]) / np.linalg.norm(rhs[b]) <= 1e-10
    bt.free()
    if m >= 20:
        A[B - 1, 7, :] = 0.0                  # a zero row: pivot 7 of the last problem is exactly 0
        bt = batched.Batch(A)
        try:
            _, status = bt.normal_solve(None, rhs)
        finally:
            bt.free()
        assert status.tolist() == [0] * (B - 1) + [1]


def test_batched_failure_is_per_problem(common):
    rng = np.random.default_rng(1)
    A = rng.random((3, 20, 30)) + np.eye(20, 30)[None]
    A[1, 5, :] = A[1, 4, :]                   # problem 1 is singular
    bt = batched.Batch(A)
    x, status = bt.normal_solve(None, rng.random((3, 20)))
    assert status.tolist() == [0, 1, 0]
    bt.free()


def _lps(B, m, n):
    sfs = [lpgen.dense_lp(m, n, seed) for seed in range(B)]
    for i, sf in enumerate(sfs):
        if i % 2:
            sf.u = np.full(n, 25.0)           # some LPs with finite upper bounds
    return sfs


@pytest.mark.parametrize("B,m,n", [(6, 24, 60), (3, 64, 160), (2, 256, 512)])
def test_batched_affine_first_iterations_match_oracle(common, B, m, n):
    """Algebra parity where it is well defined: after a fixed number of iterations (repair, optimize
    included) every LP's iterate must equal the oracle's."""
    sfs = _lps(B, m, n)
    for k in (3, 9):
        bt = batched.Batch.from_standard_forms(sfs)
        obj, x, res, iters = bt.affine_scaling(k)
        bt.free()
        for b, sf in enumerate(sfs):
            ost = oa.make_affine_state(sf.nvars, sf.ncons, sf.c_dense(), sf.A_dense, sf.b, sf.l, sf.u)
            oa.affine_scaling(ost, k)
            assert iters[b] == k
            np.testing.assert_allclose(x[b], ost.x, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("B,m,n", [(6, 24, 60), (3, 64, 160)])
def test_batched_affine_scaling_full_solve(common, B, m, n):
    """Whole solves.  Affine scaling terminates on noise-level quantities (sign of g.c ~ 1e-9, residual
    drifting across 1e-6*m), so the reference's own iteration count moves by a few steps with the BLAS
    summation order (the single-problem GPU driver and the oracle differ the same way); the objective is
    what is pinned."""
    sfs = _lps(B, m, n)
    bt = batched.Batch.from_standard_forms(sfs)
    obj, x, res, iters = bt.affine_scaling(3000)
    bt.free()
    for b, sf in enumerate(sfs):
        ost = oa.make_affine_state(sf.nvars, sf.ncons, sf.c_dense(), sf.A_dense, sf.b, sf.l, sf.u)
        oobj, ox, ores, oit = oa.affine_scaling(ost, 3000)
        assert abs(int(iters[b]) - oit) <= 5, (b, iters[b], oit)
        assert abs(obj[b] - oobj) <= 1e-6 * abs(oobj), (b, obj[b], oobj)
        assert res[b] <= 1e-6 * m
