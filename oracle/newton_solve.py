"""Oracle (test infrastructure): KKT Newton solve by block elimination to normal equations.

Restates, step by step and in the reference's operation order,
  * newton-solve.lisp:27-154          (dense A, matlisp)           -> solve_kkt_newton(..., filters=False)
  * sparse-newton-solve.lisp:30-168   (CHOLMOD sparse A, filters)  -> solve_kkt_newton(..., filters=True)
  * sparse-cholesky.lisp:409-431      solve-dense: factorize B*B' and solve system A (= LL')
  * sparse-cholesky.lisp:461-473      scale-sparse!: B = A*diag(s) (cholmod_scale, CHOLMOD_COL)
  * sparse-cholesky.lisp:506-522      solve-sparse-one-shot: status != 0 -> NIL

Block system (newton-solve.lisp:156-162):
    U dw - W dx        = e
    Z dx + L dz        = f
    A dx               = g
    A' dy + dz - dw    = h

A may be a dense ndarray (m x n) or a scipy.sparse matrix; the algebra is identical.  CHOLMOD's
supernodal LL' is played by OpenBLAS dpotrf (scipy.linalg.cho_factor) on the explicitly formed
B*B' -- for a dense A CHOLMOD itself ends up with one dense supernode and calls dpotrf.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.sparse as sp


def _matvec(A, x):
    return np.asarray(A @ x).ravel()


def _rmatvec(A, y):
    return np.asarray(A.T @ y).ravel()


def normal_matrix(A, s):
    """(A diag(s)) (A diag(s))' as a dense array.  This is the matrix CHOLMOD factorizes when
    handed the unsymmetric B = A*diag(s) (stype 0 => it works on B*B'; sparse-cholesky.lisp:408)."""
    if sp.issparse(A):
        B = A @ sp.diags(s)
        return np.asarray((B @ B.T).todense())
    B = A * s[None, :]
    return B @ B.T


def solve_dense(B, b):
    """solve-dense (sparse-cholesky.lisp:409-431): x with (B B') x = b, or None when the
    factorization reports a non-positive pivot (cholmod status != 0 -> NIL, :420-421)."""
    M = B @ B.T
    return solve_spd(M, b)


def solve_spd(M, b):
    try:
        c = scipy.linalg.cho_factor(M, lower=True, check_finite=False)
    except scipy.linalg.LinAlgError:
        return None
    if not np.all(np.isfinite(c[0].diagonal())):
        return None
    return scipy.linalg.cho_solve(c, b, check_finite=False)


def filter_U(u, w, e):
    """filter-U (sparse-newton-solve.lisp:30-38): rows with a huge slack only say w = 0."""
    big = u > 1e7
    e[big] = w[big]
    w[big] = 0.0
    u[big] = 1.0
    return u, w, e


def filter_Z(z, l, f):
    """filter-Z (sparse-newton-solve.lisp:40-45)."""
    big = l > 1e7
    f[big] = z[big]
    z[big] = 0.0
    l[big] = 1.0
    return z, l, f


def solve_kkt_newton(l, u, w, z, A, e, f, g, h, filters=False, return_intermediates=False):
    """solve-kkt-newton (newton-solve.lisp:139-154 / sparse-newton-solve.lisp:150-168).

    Inputs are slacks l = x - lo, u = hi - x, multipliers w, z, and right-hand sides e, f (n),
    g (m), h (n).  Returns (dw, dx, dy, dz), or None when the Cholesky fails.  Inputs are copied
    (the Lisp destroys them; callers there pass copies too).
    """
    l, u, w, z, e, f, g, h = (np.array(v, dtype=np.float64, copy=True) for v in (l, u, w, z, e, f, g, h))
    if filters:
        u, w, e = filter_U(u, w, e)
        z, l, f = filter_Z(z, l, f)
    # scale-U (newton-solve.lisp:27-30): w' = w/u, e' = e/u  (1/u then multiply, as map-matrix! #'/)
    inv_u = 1.0 / u
    w = inv_u * w
    e = inv_u * e
    # scale-Z (:32-33): l' = l/z, f' = f/z
    inv_z = 1.0 / z
    l = inv_z * l
    f = inv_z * f
    # clear-delta-w (:58-59): h <- e + h
    h = e + h
    # clear-delta-x-g (:61-62): g <- g - A f
    g = g - _matvec(A, f)
    # clear-delta-x-h (:64-66): wl1 = 1 + w.l ; h <- w.f + h
    wl1 = 1.0 + w * l
    h = w * f + h
    # scale-wl+1 (:92-94): d = 1/wl1 ; h <- d.h
    d = 1.0 / wl1
    h = d * h
    # clear-delta-z (:96-98): s = sqrt(l.d) ; g <- g + A (l.h)
    s = np.sqrt(l * d)
    g = g + _matvec(A, l * h)
    # solve-delta-y (:112-116 / sparse :121-126): (A diag s)(A diag s)' dy = g
    M = normal_matrix(A, s)
    dy = solve_spd(M, g)
    if dy is None:
        return None
    # solve-delta-z (:127-131): dz = h - d.(A' dy)
    dz = h - d * _rmatvec(A, dy)
    # solve-delta-x (:133-134): dx = f - dz.l
    dx = f - dz * l
    # solve-delta-w (:136-137): dw = dx.w + e
    dw = dx * w + e
    if return_intermediates:
        return dw, dx, dy, dz, {"theta": l * d, "s": s, "rhs": g, "M": M, "d": d}
    return dw, dx, dy, dz


def kkt_residuals(l, u, w, z, A, e, f, g, h, dw, dx, dy, dz, ord=2):
    """test-kkt-solve (newton-solve.lisp:166-182, 2-norm; sparse-newton-solve.lisp:180-198,
    inf-norm): residuals of the four un-reduced block rows."""
    n_ = (lambda v: np.linalg.norm(v, ord))
    return (
        n_(u * dw - w * dx - e),
        n_(z * dx + l * dz - f),
        n_(_matvec(A, dx) - g),
        n_(_rmatvec(A, dy) + dz - dw - h),
    )


# ---- the reference's own random test generators ---------------------------------------------

def random_positive_vector(rng, n):
    """random-positive-vector (newton-solve.lisp:184-185): 0.1 + 10 U(0,1)."""
    return 0.1 + 10.0 * rng.random(n)


def random_dense_case(rng, m, n):
    """test-m-n (newton-solve.lisp:187-200): A = rand(m,n) + eye(m,n); rhs U(0,1)."""
    l, u, w, z = (random_positive_vector(rng, n) for _ in range(4))
    A = rng.random((m, n)) + np.eye(m, n)
    e, f = rng.random(n), rng.random(n)
    g, h = rng.random(m), rng.random(n)
    return l, u, w, z, A, e, f, g, h


def random_sparse_matrix(rng, m, n, density=5e-2):
    """random-sparse-vector (sparse-newton-solve.lisp:228-237): entry (i,j) present when i == j
    or U(0,1) < density; value 1 + U(0,1)."""
    mask = rng.random((m, n)) < density
    idx = np.arange(min(m, n))
    mask[idx, idx] = True
    vals = 1.0 + rng.random((m, n))
    return sp.csc_matrix(np.where(mask, vals, 0.0))


def random_sparse_case(rng, m, n):
    l, u, w, z = (random_positive_vector(rng, n) for _ in range(4))
    A = random_sparse_matrix(rng, m, n)
    e, f = rng.random(n), rng.random(n)
    g, h = rng.random(m), rng.random(n)
    return l, u, w, z, A, e, f, g, h
