"""The oracle against the reference's own tests.  The reference ships no golden vectors; its tests
are randomized property checks of the four KKT block residuals (newton-solve.lisp:163-211 with the
2-norm, sparse-newton-solve.lisp:177-269 with the inf-norm; report threshold 1e-6, assert 1e-4).
We re-run them with the reference's generators, then pin the oracle's IPM behaviour in
tests/golden/ so later changes cannot drift silently."""
import json
import os

import numpy as np
import pytest

from cholesky_is_magic_b200 import lpgen
from oracle import newton_solve as ons
from oracle import pdas as opdas

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_dense_kkt_property_like_reference_test():
    """(test max) of newton-solve.lisp:202-211 with max = 20, 3 reps."""
    rng = np.random.default_rng(1234)
    worst = 0.0
    for m in range(1, 21):
        for n in range(m, 21):
            for _ in range(3):
                case = ons.random_dense_case(rng, m, n)
                out = ons.solve_kkt_newton(*case)
                assert out is not None
                worst = max(worst, max(ons.kkt_residuals(*case, *out, ord=2)))
    assert worst <= 1e-6


def test_sparse_kkt_property_like_reference_test():
    """(test max) of sparse-newton-solve.lisp:260-269 (inf-norm, filters on)."""
    rng = np.random.default_rng(4321)
    worst = 0.0
    for m in range(1, 21):
        for n in range(m, 21):
            for _ in range(3):
                case = ons.random_sparse_case(rng, m, n)
                out = ons.solve_kkt_newton(*case, filters=True)
                assert out is not None
                worst = max(worst, max(ons.kkt_residuals(*case, *out, ord=np.inf)))
    assert worst <= 1e-6


def test_theta_identity():
    """SURVEY 8a key identity: the factorized matrix is A diag(theta) A', theta = 1/(z/l + w/u)."""
    rng = np.random.default_rng(7)
    l, u, w, z, A, e, f, g, h = ons.random_dense_case(rng, 9, 17)
    *_, inter = ons.solve_kkt_newton(l, u, w, z, A, e, f, g, h, return_intermediates=True)
    np.testing.assert_allclose(inter["theta"], 1.0 / (z / l + w / u), rtol=1e-13)
    np.testing.assert_allclose(inter["M"], (A * inter["theta"]) @ A.T, rtol=1e-12)


def test_solve_dense_reports_failure_as_none():
    B = np.zeros((3, 5))
    assert ons.solve_dense(B, np.ones(3)) is None   # singular -> NIL (sparse-cholesky.lisp:420-421)


def test_make_pdas_init_rules():
    # primal-dual-affine-scaling.lisp:88-118
    cvec = np.array([0.0, -2.0, 3.0, 1.0, 1.0])
    lo = np.array([0.0, -np.inf, -np.inf, 2.0, 1.0])
    hi = np.array([np.inf, np.inf, 5.0, 4.0, 1.0])
    A = np.ones((1, 5))
    st = opdas.make_pdas(5, 1, cvec, A, np.ones(1), lo, hi, scale=False)
    np.testing.assert_allclose(st.x, [1.0, 0.0, 5.0 - 1.5, 3.0, 1.0])
    np.testing.assert_allclose(st.z, [1, 1, 4, 2, 2])
    np.testing.assert_allclose(st.w, [1, 3, 1, 1, 1])
    assert st.l[4] == 1.0 - 5e-7 and st.u[4] == 1.0 + 5e7     # near-fixed widening, sic
    assert st.u[0] == 1e8 and st.l[1] == -1e8                    # clamp


def test_step_rules():
    l = np.array([1.0, 2.0, 3.0]); u = np.array([4.0, 5.0, 6.0])
    dx = np.array([2.0, -10.0, 0.0])
    # subtractive update: dx>0 moves toward the lower bound (l/dx), dx<0 toward the upper (u/-dx)
    assert opdas.box_step_vec(l, u, dx) == pytest.approx(0.5)
    assert opdas.box_step(l, u, dx) == pytest.approx(0.5)
    assert opdas.pos_step(np.array([1.0, 2.0]), np.array([-1.0, 4.0])) == pytest.approx(0.5)
    assert opdas.pos_step(np.array([1.0]), np.array([-1.0])) == np.inf
    assert opdas.max_step(np.zeros(2), np.ones(2), np.full(2, 3.0), np.array([-2.0, 1.0])) == pytest.approx(0.5)


@pytest.mark.parametrize("name", ["pdas_dense_m20_n50_seed0", "pdas_dense_m200_n500_seed0"])
def test_pdas_oracle_matches_golden(name):
    """Golden fixtures were produced by tests/golden/make_golden.py from this oracle."""
    path = os.path.join(GOLDEN, name + ".json")
    gold = json.load(open(path))
    sf = lpgen.dense_lp(gold["m"], gold["n"], gold["seed"])
    st = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), sf.A_dense, sf.b, sf.l, sf.u)
    obj, gap, iters = opdas.pdas(st, 500)
    assert iters == gold["iterations"]
    assert obj == pytest.approx(gold["dobj"], rel=1e-9)
    assert [e["branch"] for e in st.log] == gold["branches"]
    np.testing.assert_allclose([e["gap"] for e in st.log], gold["gaps"], rtol=1e-6, atol=1e-12)
