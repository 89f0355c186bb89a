"""Known-answer anchors for the oracle (and, on a GPU, for the library): the restated reference
algorithms must reach the optimum an INDEPENDENT LP solver (HiGHS through scipy.optimize.linprog) finds.
The reference ships no golden vectors, so this is the pin that does not depend on our
own restatement: a wrong sign in the KKT elimination, in the initial point, the step rule or the stop rule
would not end at the true optimum within the reference's own tolerance (relative gap < 1e-4,
primal-dual-affine-scaling.lisp:394)."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from cholesky_is_magic_b200 import lpgen, read_mps
from oracle import affine_scaling as oa
from oracle import pdas as opdas


def highs(sf, A):
    bounds = [(lo if np.isfinite(lo) else None, hi if np.isfinite(hi) else None) for lo, hi in zip(sf.l, sf.u)]
    r = linprog(sf.c_dense(), A_eq=A, b_eq=sf.b, bounds=bounds, method="highs")
    assert r.status == 0, r.message
    return r.fun, r.x


CASES = [("dense", 20, 50, None), ("dense", 64, 160, 20.0), ("dense", 200, 500, None), ("sparse", 150, 400, None),
         ("sparse", 400, 1000, 30.0)]


def make_lp(kind, m, n, ub):
    if kind == "dense":
        sf = lpgen.dense_lp(m, n, 0)
        A = sf.A_dense
    else:
        sf = lpgen.sparse_lp(m, n, nnz_per_col=5, bandwidth=20, seed=2)
        A = sp.csc_matrix((sf.A.value, (sf.A.row, sf.A.col)), shape=(m, n))
    if ub is not None:
        sf.u = np.full(n, ub)
    return sf, A


@pytest.mark.parametrize("kind,m,n,ub", CASES)
def test_oracle_pdas_reaches_the_highs_optimum(kind, m, n, ub):
    sf, A = make_lp(kind, m, n, ub)
    fun, _ = highs(sf, A)
    st = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), A, sf.b, sf.l, sf.u)
    obj, gap, it = opdas.pdas(st, 400)
    assert gap < 1e-4
    assert abs(obj - fun) <= 2e-4 * abs(fun)          # the stop rule is a 1e-4 relative duality gap


@pytest.mark.parametrize("kind,m,n,ub", CASES[:4])
def test_oracle_affine_scaling_reaches_the_highs_optimum(kind, m, n, ub):
    sf, A = make_lp(kind, m, n, ub)
    fun, _ = highs(sf, A)
    st = oa.make_affine_state(sf.nvars, sf.ncons, sf.c_dense(), A, sf.b, sf.l, sf.u)
    obj, x, res, it = oa.affine_scaling(st, 3000)
    assert np.linalg.norm(res) <= 1e-6 * m
    assert 0 <= obj - fun <= 1e-5 * abs(fun)           # a primal method: approaches the optimum from above


TEXTBOOK_MPS = """NAME          TEXTBOOK
ROWS
 N  COST
 L  LIM1
 L  LIM2
 L  LIM3
COLUMNS
    X1        COST            -3.0   LIM1             1.0
    X1        LIM3             3.0
    X2        COST            -5.0   LIM2             2.0
    X2        LIM3             2.0
RHS
    RHS       LIM1             4.0   LIM2            12.0
    RHS       LIM3            18.0
ENDATA
"""


def test_mps_to_standard_form_known_answer(tmp_path):
    """max 3 x1 + 5 x2, x1 <= 4, 2 x2 <= 12, 3 x1 + 2 x2 <= 18, x >= 0: optimum 36 at (2, 6) -- through
    read-mps, to-standard-form (slack columns) and the oracle's PDAS."""
    p = tmp_path / "textbook.mps"
    p.write_text(TEXTBOOK_MPS)
    sf = read_mps.to_standard_form(read_mps.read_mps_file(p))
    A = sp.csc_matrix((sf.A.value, (sf.A.row, sf.A.col)), shape=(sf.ncons, sf.nvars))
    st = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), A, sf.b, sf.l, sf.u)
    obj, gap, it = opdas.pdas(st, 400)
    assert abs(obj - (-36.0)) <= 1e-3 * 36.0
    np.testing.assert_allclose(st.x[:2], [2.0, 6.0], atol=2e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,m,n,ub", CASES)
def test_gpu_pdas_reaches_the_highs_optimum(common, kind, m, n, ub):
    from cholesky_is_magic_b200 import pdas
    sf, A = make_lp(kind, m, n, ub)
    fun, _ = highs(sf, A)
    obj, gap, it = pdas.pdas(pdas.make_pdas(sf), 400, native_loop=True)
    assert gap < 1e-4 and abs(obj - fun) <= 2e-4 * abs(fun)
