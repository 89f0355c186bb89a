"""GPU parity tests for the dense hot path, called through the C ABI (ctypes) and checked against
the CPU oracle on the same seeded inputs.  Gates:
   formation   |M - M_ref|_F / |M_ref|_F <= 1e-13
   Cholesky    |L L' - M|_F / |M|_F     <= 1e-12
   KKT Newton  four block residuals <= 1e-6 (the reference's own threshold), directions vs oracle
   PDAS        identical iteration count, objective within 1e-9 relative
"""
import json
import os

import numpy as np
import pytest

from cholesky_is_magic_b200 import lpgen, nes, newton_solve, pdas, sparse_cholesky
from oracle import newton_solve as ons
from oracle import pdas as opdas

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


SHAPES = [(1, 1), (1, 7), (3, 3), (17, 40), (128, 128), (129, 300), (200, 500), (257, 513), (384, 1000)]


@pytest.mark.parametrize("m,n", SHAPES)
def test_formation_matches_oracle(common, m, n):
    rng = np.random.default_rng(m * 1000 + n)
    A = rng.random((m, n)) + np.eye(m, n)
    s = np.sqrt(0.1 + 10 * rng.random(n))
    Ad = nes.Matrix.from_dense(common, A)
    Ad.scale(s)
    M = Ad.normal_matrix()
    Ad.free()
    assert relerr(M, ons.normal_matrix(A, s)) <= 1e-13


def test_formation_unscaled_and_wide_theta_range(common):
    rng = np.random.default_rng(5)
    m, n = 150, 333
    A = rng.random((m, n)) + np.eye(m, n)
    Ad = nes.Matrix.from_dense(common, A)
    assert relerr(Ad.normal_matrix(), A @ A.T) <= 1e-13
    s = 10.0 ** rng.uniform(-6, 6, n)       # late-IPM like spread of theta = s^2: 24 decades
    Ad.scale(s)
    assert relerr(Ad.normal_matrix(), ons.normal_matrix(A, s)) <= 1e-13
    Ad.unscale()
    assert relerr(Ad.normal_matrix(), A @ A.T) <= 1e-13
    Ad.free()


@pytest.mark.parametrize("m,n", SHAPES + [(640, 700), (1000, 1500)])
def test_cholesky_residual_and_solve(common, m, n):
    rng = np.random.default_rng(m + 7 * n)
    A = rng.random((m, n)) + np.eye(m, n)
    s = np.sqrt(0.1 + 10 * rng.random(n))
    b = rng.random(m)
    Ad = nes.Matrix.from_dense(common, A)
    Ad.scale(s)
    L = nes.Factor(common, Ad)
    assert L.factorize(Ad)
    Lh = L.to_dense()
    M = ons.normal_matrix(A, s)
    assert relerr(Lh @ Lh.T, M) <= 1e-12                 # the north-star gate
    x = L.solve(b)
    xo = ons.solve_spd(M, b)
    assert np.linalg.norm(M @ x - b) / np.linalg.norm(b) <= 1e-10
    assert relerr(x, xo) <= 1e-8 * max(1.0, np.linalg.cond(M) * 1e-8)
    # analytic counters printed by the reference after analyze (affine-scaling.lisp:273-279)
    assert common.lnz == m * (m + 1) / 2 and common.anz == m * (m + 1) / 2
    L.free()
    Ad.free()


@pytest.mark.parametrize("m,n,defer", [(1536, 2048, 640), (1400, 1700, 512), (2048, 2304, 4096)])
def test_deferred_formation_matches_upfront_formation(common, m, n, defer, monkeypatch):
    """Opt-in path of dense_chol.cu (NES_CHOL_DEFER): the last columns of M are formed strip by strip
    inside the factorization (k-split tiles, accumulated with the trailing updates).  Same gates as the
    default path, and the two factors agree to rounding."""
    rng = np.random.default_rng(m + n)
    A = rng.random((m, n)) + np.eye(m, n)
    s = np.sqrt(0.1 + 10 * rng.random(n))
    b = rng.random(m)
    M = ons.normal_matrix(A, s)
    Ad = nes.Matrix.from_dense(common, A)
    Ad.scale(s)
    out = {}
    for mode, env in (("plain", {"NES_CHOL_DEFER": "0"}),
                      ("deferred", {"NES_CHOL_DEFER": str(defer), "NES_CHOL_DEFER_MIN_M": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        L = nes.Factor(common, Ad)              # the plan is made at the first factorization
        assert L.factorize(Ad)
        deferred_cols = round(max(0.0, m * m - common.form_flops / n) ** 0.5)
        assert (deferred_cols > 0) == (mode == "deferred"), (mode, deferred_cols)
        Lh = L.to_dense()
        assert relerr(Lh @ Lh.T, M) <= 1e-12
        x = L.solve(b)
        assert np.linalg.norm(M @ x - b) / np.linalg.norm(b) <= 1e-10
        assert L.factorize(Ad)                  # bitwise reproducible (fixed summation order)
        np.testing.assert_array_equal(L.to_dense(), Lh)
        out[mode] = Lh
        L.free()
    assert relerr(out["deferred"], out["plain"]) <= 1e-11
    Ad.free()


def test_deferred_formation_reports_a_failed_pivot(common, monkeypatch):
    monkeypatch.setenv("NES_CHOL_DEFER", "512")
    monkeypatch.setenv("NES_CHOL_DEFER_MIN_M", "0")
    rng = np.random.default_rng(9)
    B = rng.random((1300, 1250))                # m > n: B B' singular, the pivot fails in the deferred part
    Ad = nes.Matrix.from_dense(common, B)
    L = nes.Factor(common, Ad)
    L.factorize(Ad)
    assert common.status == nes.NES_NOT_POSDEF and 1240 <= common.minor <= 1300
    L.free()
    Ad.free()


def test_solve_dense_like_reference(common):
    """solve-dense (sparse-cholesky.lisp:409-431)."""
    rng = np.random.default_rng(11)
    with_A = rng.random((60, 140)) + np.eye(60, 140)
    b = rng.random(60)
    x = sparse_cholesky.solve_dense(with_A, b)
    assert relerr(x, ons.solve_dense(with_A, b)) <= 1e-9


def test_not_positive_definite_is_reported_not_masked(common):
    """status = 1 (CHOLMOD_NOT_POSDEF) -> NIL (sparse-cholesky.lisp:418-421)."""
    B = np.zeros((5, 9))
    B[0, 0] = 1.0
    assert sparse_cholesky.solve_dense(B, np.ones(5)) is None
    assert common.status == nes.NES_NOT_POSDEF
    assert common.minor == 1
    # rank deficient inside a later block
    rng = np.random.default_rng(3)
    B = rng.random((200, 150))              # m > n: B B' singular
    assert sparse_cholesky.solve_dense(B, np.ones(200)) is None
    assert 140 <= common.minor <= 200


def test_dbound_clamps_the_diagonal(common):
    B = np.zeros((4, 6))
    B[0, 0] = 1.0
    common.set("dbound", 0.5)
    x = sparse_cholesky.solve_dense(B, np.ones(4))
    common.set("dbound", 0.0)
    assert x is not None and np.all(np.isfinite(x))
    np.testing.assert_allclose(x, [1.0, 4.0, 4.0, 4.0])   # L = diag(1, .5, .5, .5)


@pytest.mark.parametrize("m,n", [(1, 1), (5, 9), (200, 500), (300, 257 + 300)])
def test_sdmult_matches_numpy(common, m, n):
    rng = np.random.default_rng(m + n)
    A = rng.standard_normal((m, n))
    x, y = rng.standard_normal(n), rng.standard_normal(m)
    Ad = nes.Matrix.from_dense(common, A)
    np.testing.assert_allclose(Ad.sdmult(x), A @ x, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Ad.sdmult(y, transpose=True), A.T @ y, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Ad.sdmult(x, y=y, alpha=-1.0), y - A @ x, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Ad.sdmult(y, transpose=True, y=x, alpha=-1.0, beta=2.0),
                               2 * x - A.T @ y, rtol=1e-12, atol=1e-12)
    s = 0.5 + rng.random(n)
    Ad.scale(s)
    np.testing.assert_allclose(Ad.sdmult(x), (A * s) @ x, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(Ad.sdmult(y, transpose=True), (A * s).T @ y, rtol=1e-12, atol=1e-12)
    Ad.free()


def test_kkt_newton_property_like_reference_test(common):
    """(test max) of newton-solve.lisp:202-211: residuals of the un-reduced block system, 2-norm,
    report threshold 1e-6; data from the reference's generators."""
    rng = np.random.default_rng(2024)
    worst = 0.0
    for m in list(range(1, 14)) + [40, 130]:
        for n in (m, m + 1, 2 * m + 3):
            l, u, w, z, A, e, f, g, h = ons.random_dense_case(rng, m, n)
            Ad = nes.Matrix.from_dense(common, A)
            res = newton_solve.test_kkt_solve(l, u, w, z, Ad, e, f, g, h, A)
            Ad.free()
            worst = max(worst, max(res))
    assert worst <= 1e-6


@pytest.mark.parametrize("name", ["kkt_dense_m12_n30_seed5", "kkt_dense_m130_n300_seed6"])
def test_kkt_newton_matches_golden(common, name):
    gz = np.load(os.path.join(GOLDEN, name + ".npz"))
    Ad = nes.Matrix.from_dense(common, gz["A"])
    dw, dx, dy, dz = newton_solve.solve_kkt_newton(gz["l"], gz["u"], gz["w"], gz["z"], Ad, gz["e"], gz["f"],
                                                   gz["g"], gz["h"])
    Ad.free()
    for got, key in ((dw, "dw"), (dx, "dx"), (dy, "dy"), (dz, "dz")):
        assert relerr(got, gz[key]) <= 1e-9, key


def test_kkt_newton_filters(common):
    rng = np.random.default_rng(99)
    l, u, w, z, A, e, f, g, h = ons.random_dense_case(rng, 30, 70)
    u[::3] = 1e8                             # filter-U fires (sparse-newton-solve.lisp:30-38)
    Ad = nes.Matrix.from_dense(common, A)
    got = newton_solve.solve_kkt_newton(l, u, w, z, Ad, e, f, g, h, filters=True)
    want = ons.solve_kkt_newton(l, u, w, z, A, e, f, g, h, filters=True)
    Ad.free()
    for a, b in zip(got, want):
        assert relerr(a, b) <= 1e-9


@pytest.mark.parametrize("m,n", [(20, 50), (200, 500)])
def test_pdas_matches_oracle_and_golden(common, m, n):
    """BASELINE config 1 (m=200, n=500): identical iteration count, objective within 1e-9."""
    sf = lpgen.dense_lp(m, n, 0)
    ost = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), sf.A_dense, sf.b, sf.l, sf.u)
    oobj, ogap, oit = opdas.pdas(ost, 500)
    st = pdas.make_pdas(sf)
    obj, gap, it = pdas.pdas(st, 500)
    assert it == oit
    assert abs(obj - oobj) <= 1e-9 * abs(oobj)
    assert [e["branch"] for e in st.log] == [e["branch"] for e in ost.log]
    # per-iteration trajectory (diagnostic; the gates are the iteration count and the objective):
    # intermediate iterates amplify summation-order differences through the conditioning of M, and
    # dobj passes through zero on its way up from -n*1e8, so compare on the scale of the largest value
    for key in ("pobj", "dobj"):
        got, want = np.array([e[key] for e in st.log]), np.array([e[key] for e in ost.log])
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= 1e-4, key
    np.testing.assert_allclose([e["gap"] for e in st.log], [e["gap"] for e in ost.log], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose([e["step"] for e in st.log if "step" in e],
                               [e["step"] for e in ost.log if "step" in e], rtol=1e-4)
    np.testing.assert_allclose(st.final["x"], ost.x, rtol=1e-6, atol=1e-9)
    gold = json.load(open(os.path.join(GOLDEN, f"pdas_dense_m{m}_n{n}_seed0.json")))
    assert it == gold["iterations"] and abs(obj - gold["dobj"]) <= 1e-9 * abs(gold["dobj"])


def test_pdas_native_loop_equals_stepwise(common):
    sf = lpgen.dense_lp(60, 150, 3)
    a = pdas.pdas(pdas.make_pdas(sf), 500)
    b = pdas.pdas(pdas.make_pdas(sf), 500, native_loop=True)
    assert a[2] == b[2] and a[0] == b[0] and a[1] == b[1]


def test_generated_matrix_equals_host_generator(common):
    m, n, seed = 70, 190, 42
    Ad = nes.Matrix.generate_dense(common, m, n, seed)
    A = lpgen.dense_matrix(m, n, seed)
    x = np.arange(n, dtype=np.float64)
    np.testing.assert_allclose(Ad.sdmult(np.eye(n)[3]), A[:, 3], rtol=0, atol=0)
    np.testing.assert_allclose(Ad.sdmult(x), A @ x, rtol=1e-13)
    Ad.free()
