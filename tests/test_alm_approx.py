"""ALM + APPROX (alm-approx.lisp).  CPU: the oracle's `alm` must reach the HiGHS optimum
with |Ax - b|_inf <= 1e-5.  GPU: the inner solver (variant 1) against the oracle on a subproblem, and the
whole `alm` against the oracle and HiGHS."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from cholesky_is_magic_b200 import lpgen
from oracle import alm_approx as oalm


def lp(kind, m, n, seed, ub=None):
    sf = lpgen.sparse_lp(m, n, nnz_per_col=4, bandwidth=8, seed=seed) if kind == "sparse" else lpgen.dense_lp(m, n, seed)
    if ub is not None:
        sf.u = np.full(n, ub)
    A = sp.csr_matrix((sf.A.value, (sf.A.row, sf.A.col)), shape=(m, n)) if sf.A is not None else sp.csr_matrix(sf.A_dense)
    return sf, A


def highs(sf, A):
    bounds = [(lo if np.isfinite(lo) else None, hi if np.isfinite(hi) else None) for lo, hi in zip(sf.l, sf.u)]
    r = linprog(sf.c_dense(), A_eq=A, b_eq=sf.b, bounds=bounds, method="highs")
    assert r.status == 0
    return r.fun


@pytest.mark.parametrize("kind,m,n,ub", [("dense", 8, 20, None), ("dense", 12, 30, 15.0), ("sparse", 30, 70, None)])
def test_oracle_alm_reaches_the_highs_optimum(kind, m, n, ub):
    sf, A = lp(kind, m, n, 1, ub)
    st = oalm.make_alm(A, sf.b, sf.c_dense(), sf.l, sf.u, list(sf.type))
    outer, inner, v, pg, value, x = oalm.alm(st, maxiter=200)
    fun = highs(sf, A)
    assert v <= 1e-5 and pg <= 1e-5
    assert abs(float(sf.c_dense() @ x) - fun) <= 1e-4 * abs(fun)
    assert abs(value - fun) <= 1e-5 * abs(fun)            # the dual value converges faster
    assert np.all(x >= sf.l - 1e-12) and np.all(x <= sf.u + 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,m,n,ub", [("dense", 12, 30, 15.0), ("sparse", 60, 150, None)])
def test_gpu_inner_solver_matches_oracle_on_a_subproblem(common, kind, m, n, ub):
    from cholesky_is_magic_b200 import alm_approx as galm
    sf, A = lp(kind, m, n, 2, ub)
    rng = np.random.default_rng(0)
    lam = rng.standard_normal(m)
    st = galm.make_alm(sf, mu=7.0, multipliers=lam)
    try:
        for k in (3, 40):
            sub = oalm.make_alm_subproblem(A, sf.b, sf.c_dense(), sf.l, sf.u, lam, 7.0)
            oz, opg, oit, orest = oalm.approx(sub, k, None, 1e-30)       # accuracy never reached: exactly k steps
            st.multipliers = lam.copy(); st.mu = 7.0
            x, viol, pg, value, it = galm.alm_iteration2(st, None, precision=1e-30, max_inner=k)
            assert it == oit == k
            np.testing.assert_allclose(x, oz, rtol=1e-9, atol=1e-11)
            assert abs(pg - opg) <= 1e-8 * max(opg, 1e-12)
            assert abs(value - oalm.dual_value(sub, oz)) <= 1e-9 * abs(value)
    finally:
        st.free()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,m,n,ub", [("dense", 12, 30, 15.0), ("sparse", 30, 70, None), ("sparse", 200, 500, 10.0)])
def test_gpu_alm_matches_oracle_and_highs(common, kind, m, n, ub):
    from cholesky_is_magic_b200 import alm_approx as galm
    sf, A = lp(kind, m, n, 1, ub)
    fun = highs(sf, A)
    st = galm.make_alm(sf)
    try:
        outer, inner, v, pg, value, x = galm.alm(st, maxiter=300)
    finally:
        st.free()
    assert v <= 1e-5 and pg <= 1e-5
    assert abs(value - fun) <= 1e-5 * abs(fun)
    ost = oalm.make_alm(A, sf.b, sf.c_dense(), sf.l, sf.u, list(sf.type))
    oouter, oinner, ov, opg, ovalue, ox = oalm.alm(ost, maxiter=300)
    # the inner stop (|pg| < accuracy) is reached at slightly different iterations when rounding differs, so
    # counts agree closely but not exactly; objective and multipliers are what is pinned
    assert abs(outer - oouter) <= 2 and abs(inner - oinner) <= 0.1 * oinner + 20
    assert abs(value - ovalue) <= 1e-6 * abs(ovalue)


def test_oracle_other_outer_loops_make_progress():
    """alm-iteration (minor/major rule), aalm and adcd-iteration drive |Ax - b| down on a small LP."""
    sf, A = lp("dense", 8, 20, 1)
    fun = highs(sf, A)
    st = oalm.make_alm(A, sf.b, sf.c_dense(), sf.l, sf.u, list(sf.type))
    x = None
    for _ in range(12):
        x, viol, value, kind = oalm.alm_iteration(st, x)
        assert kind in ("minor", "major")
    assert np.linalg.norm(viol) <= 1e-3 and abs(value - fun) <= 1e-3 * abs(fun)
    st = oalm.make_alm(A, sf.b, sf.c_dense(), sf.l, sf.u, list(sf.type))
    outer, inner, v, pg, z, x = oalm.aalm(st, maxiter=40)
    assert v <= 1e-3
    st = oalm.make_alm(A, sf.b, sf.c_dense(), sf.l, sf.u, list(sf.type))
    x, done = None, False
    for _ in range(200):
        x, viol, done = oalm.adcd_iteration(st, x)
        if done:
            break
    assert done and np.linalg.norm(viol) < 1e-2
    assert oalm.next_extrapolation(1.0) == 0.5 * (1 + 5 ** 0.5)


@pytest.mark.gpu
def test_gpu_other_outer_loops_match_oracle(common):
    """Same sequences of outer iterations as the oracle (kinds, mu, multipliers) for alm-iteration and
    adcd-iteration; aalm ends within the same tolerance."""
    from cholesky_is_magic_b200 import alm_approx as galm
    sf, A = lp("sparse", 30, 70, 1)
    ost = oalm.make_alm(A, sf.b, sf.c_dense(), sf.l, sf.u, list(sf.type))
    st = galm.make_alm(sf)
    try:
        x = ox = None
        for k in range(6):
            x, viol, value, kind = galm.alm_iteration(st, x)
            ox, oviol, ovalue, okind = oalm.alm_iteration(ost, ox)
            assert kind == okind and abs(st.mu - ost.mu) <= 1e-12 * ost.mu
            assert abs(value - ovalue) <= 1e-4 * abs(ovalue) + 1e-6
            np.testing.assert_allclose(st.multipliers, ost.multipliers, rtol=1e-3, atol=1e-4)
    finally:
        st.free()
    ost = oalm.make_alm(A, sf.b, sf.c_dense(), sf.l, sf.u, list(sf.type))
    st = galm.make_alm(sf)
    try:
        x = ox = None
        for k in range(8):
            x, viol, done = galm.adcd_iteration(st, x)
            ox, oviol, odone = oalm.adcd_iteration(ost, ox)
            assert done == odone and abs(st.mu - ost.mu) <= 1e-12 * ost.mu
            np.testing.assert_allclose(x, ox, rtol=1e-6, atol=1e-8)
            if done:
                break
    finally:
        st.free()
    st = galm.make_alm(sf)
    try:
        outer, inner, v, pg, z, x = galm.aalm(st, maxiter=60)
        assert v <= 1e-3
    finally:
        st.free()
