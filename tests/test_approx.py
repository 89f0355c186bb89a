"""First-order solver of approx.lisp.  CPU: the oracle restatement against hand-checked
properties (gradient by finite differences, the KKT point of the LP is a zero of the penalised objective,
convergence toward the HiGHS optimum).  GPU: parity of make-approx (nu, scales), value-&-gradient and the
iterates of `approx` against the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from cholesky_is_magic_b200 import lpgen
from oracle import approx as oap


def small_lp(m=12, n=30, seed=3, ub=None, sparse=False):
    if sparse:
        sf = lpgen.sparse_lp(m, n, nnz_per_col=4, bandwidth=8, seed=seed)
    else:
        sf = lpgen.dense_lp(m, n, seed)
    if ub is not None:
        sf.u = np.full(n, ub)
    return sf


def triplets_of(sf):
    if sf.A is not None:
        return list(zip(sf.A.row.tolist(), sf.A.col.tolist(), sf.A.value.tolist()))
    r, c = np.nonzero(sf.A_dense)
    return list(zip(r.tolist(), c.tolist(), sf.A_dense[r, c].tolist()))


def oracle_state(sf, **kw):
    return oap.make_approx(sf.nvars, sf.ncons, list(sf.c), triplets_of(sf), sf.b, list(sf.type), sf.l, sf.u, **kw)


def test_oracle_gradient_matches_finite_differences():
    sf = small_lp(ub=15.0)
    st = oracle_state(sf, complementarity=False, l1_penalty=0.01)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(st.nvars)
    f0, g, mx = oap.value_and_gradient(st, x)
    assert mx > 0 and len(g) == 3 * sf.nvars + sf.ncons
    for j in rng.choice(st.nvars, 12, replace=False):
        e = np.zeros(st.nvars); e[j] = 1e-6
        fd = (oap.value_and_gradient(st, x + e)[0] - oap.value_and_gradient(st, x - e)[0]) / 2e-6
        assert abs(fd - g[j]) <= 1e-6 * max(1.0, abs(g[j]))


def test_oracle_complementarity_terms_follow_the_reference_formula():
    """(x - x0)+ (y - y0)+ with the reference's gradient (approx.lisp:310-326): the CLAMPED other factor is
    added even where this factor was clamped to zero, so it is not the derivative there -- restated as is."""
    sf = small_lp(ub=15.0)
    st0 = oracle_state(sf, complementarity=False)
    st1 = oracle_state(sf, complementarity=True)
    nv, nc = sf.nvars, sf.ncons
    rng = np.random.default_rng(1)
    x = rng.standard_normal(st0.nvars) * 5
    v0, g0, _ = oap.value_and_gradient(st0, x)
    v1, g1, _ = oap.value_and_gradient(st1, x)
    xs, z, w = x[:nv], x[nv + nc: 2 * nv + nc], x[2 * nv + nc:]
    xz, zz = np.maximum(xs - sf.l, 0), np.maximum(z, 0)
    xw, ww = np.maximum(sf.u - xs, 0), np.maximum(w, 0)
    assert abs((v1 - v0) - (xz @ zz + xw @ ww)) <= 1e-10 * abs(v1)
    np.testing.assert_allclose(g1[:nv] - g0[:nv], zz - ww, atol=1e-12)
    np.testing.assert_allclose(g1[nv + nc: 2 * nv + nc] - g0[nv + nc: 2 * nv + nc], xz, atol=1e-12)
    np.testing.assert_allclose(g1[2 * nv + nc:] - g0[2 * nv + nc:], xw, atol=1e-12)


def test_oracle_kkt_point_is_a_zero_and_approx_converges_to_it():
    sf = small_lp(m=8, n=20, seed=1)
    A = sf.A_dense
    r = linprog(sf.c_dense(), A_eq=A, b_eq=sf.b, bounds=[(0, None)] * sf.nvars, method="highs")
    y = r.eqlin.marginals
    zdual = sf.c_dense() - A.T @ y                       # reduced costs >= 0 (lower bounds at 0, no upper bounds)
    st = oracle_state(sf)
    v = np.concatenate([r.x, y, np.maximum(zdual, 0.0), np.zeros(sf.nvars)])
    val, g, mx = oap.value_and_gradient(st, v)
    assert val <= 1e-12                                  # primal, dual and gap rows are all satisfied
    assert np.linalg.norm(oap.project_gradient(st, v, g)) <= 1e-6
    z, it, restarts = oap.approx(st, 4000)
    val_end = oap.value_and_gradient(st, z)[0]
    assert val_end <= 1e-3 * oap.value_and_gradient(st, np.clip(np.zeros(st.nvars), st.l, st.u))[0]
    assert st.log and st.log[0][0] == 1 and restarts >= 0


def test_oracle_bounds_layout_and_nu():
    sf = small_lp(ub=15.0)
    st = oracle_state(sf)
    nv, nc = sf.nvars, sf.ncons
    assert np.all(st.l[nv + nc:] == 0.0) and np.all(st.u[nv + nc:] == np.inf)     # z, w >= 0, both active
    assert np.all(st.l[nv: nv + nc] == -np.inf)                                    # y free (equality rows)
    sf2 = small_lp()                                                               # no upper bounds: w fixed at 0
    st2 = oracle_state(sf2)
    assert np.all(st2.u[2 * nv + nc:] == 0.0)
    quads = [c for c in st.constraints if isinstance(c, oap.Quadratic)]
    assert len(quads) == nc + nv + 1
    K = sp.lil_matrix((len(quads), st.nvars))
    for i, q in enumerate(quads):
        K[i, q.indices] = q.coefs * q.scale
    K = sp.csc_matrix(K)
    beta = np.array([q.beta for q in quads])
    nu = np.asarray(K.multiply(K).T @ beta).ravel()
    np.testing.assert_allclose(st.nu, nu, rtol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,comp,ub", [("dense", False, None), ("dense", True, 15.0), ("sparse", False, 12.0),
                                          ("sparse", True, None)])
def test_gpu_make_approx_value_gradient_and_iterates_match_oracle(common, kind, comp, ub):
    from cholesky_is_magic_b200 import approx as gap
    sf = small_lp(m=40, n=90, seed=2, ub=ub, sparse=(kind == "sparse"))
    ost = oracle_state(sf, complementarity=comp, l1_penalty=0.02)
    st = gap.make_approx(sf, complementarity=comp, l1_penalty=0.02)
    try:
        np.testing.assert_allclose(st.nu, ost.nu, rtol=1e-12, atol=1e-300)
        np.testing.assert_array_equal(st.l, ost.l)
        np.testing.assert_array_equal(st.u, ost.u)
        oscale = np.array([c.scale for c in ost.constraints if isinstance(c, oap.Quadratic)])
        np.testing.assert_allclose(st.scale, oscale, rtol=1e-13)
        rng = np.random.default_rng(7)
        x = rng.standard_normal(st.nvars)
        val, g, mx = gap.value_and_gradient(st, x)
        oval, og, omx = oap.value_and_gradient(ost, x)
        assert abs(val - oval) <= 1e-12 * abs(oval) and abs(mx - omx) <= 1e-12 * abs(omx)
        np.testing.assert_allclose(g, og, rtol=1e-11, atol=1e-12)
        for k in (1, 7, 60):
            z, it, restarts, stats = gap.approx(st, k)
            ost.log.clear()
            oz, oit, orestarts = oap.approx(ost, k)
            assert it == oit and restarts == orestarts
            np.testing.assert_allclose(z, oz, rtol=1e-9, atol=1e-11)
            n_it, gnorm, pg, omax, oval_end, _ = ost.log[-1]
            assert n_it == it
            assert abs(stats[0] - gnorm) <= 1e-9 * gnorm and abs(stats[1] - pg) <= 1e-8 * max(pg, 1e-12)
            assert abs(stats[3] - oval_end) <= 1e-9 * abs(oval_end)
    finally:
        st.free()


@pytest.mark.gpu
def test_gpu_approx_reduces_the_penalised_objective(common):
    from cholesky_is_magic_b200 import approx as gap
    sf = small_lp(m=300, n=700, seed=5, sparse=True)
    st = gap.make_approx(sf)
    try:
        v0 = gap.value_and_gradient(st, np.clip(np.zeros(st.nvars), st.l, st.u))[0]
        z, it, restarts, stats = gap.approx(st, 3000)
        assert stats[3] <= 1e-2 * v0
        z2, it2, r2, stats2 = gap.approx(st, 3000)            # bitwise reproducible
        np.testing.assert_array_equal(z, z2)
    finally:
        st.free()
