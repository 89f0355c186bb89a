"""Oracle (test infrastructure): primal affine scaling, restating affine-scaling.lisp function by
function in NumPy (parity unpinned by the reference, see oracle/__init__.py).

    min c'x  s.t.  Ax = b,  l <= x <= u        (bounds NOT clamped here, affine-scaling.lisp:52-90)

CHOLMOD's analyze-once + numeric refactorization (solve-sparse-recycle) is played by a dense Cholesky
of (A diag(s))(A diag(s))' per iteration.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import newton_solve as ns

MAX_SLACK = 1e8   # *max-slack*, affine-scaling.lisp:118
GAMMA = 0.9       # *gamma*, :135


def _mv(A, x):
    return np.asarray(A @ x).ravel()


def _rmv(A, y):
    return np.asarray(A.T @ y).ravel()


@dataclass
class State:
    """affine-scaling-state (:1-10)"""
    nvars: int
    ncons: int
    x: np.ndarray
    c: np.ndarray
    A: object
    b: np.ndarray
    l: np.ndarray
    u: np.ndarray
    log: list = field(default_factory=list)


def make_affine_state(nvars, ncons, cvec, A, b, sf_l, sf_u):
    """make-affine-state (:52-90): near-fixed widening, initial x (thresholds 1e10; the lower-bounded
    case uses 1 + |l| * 1.0, :74-75)."""
    l = np.array(sf_l, dtype=np.float64, copy=True)
    u = np.array(sf_u, dtype=np.float64, copy=True)
    x = np.zeros(nvars)
    for i in range(nvars):
        if u[i] - l[i] < 1e-6:
            l[i] -= 5e-7
            u[i] += 5e7
        lo, hi = l[i], u[i]
        delta = hi - lo
        if lo < -1e10 and hi > 1e10:
            x[i] = 0.0
        elif lo < -1e10:
            x[i] = hi - min(delta / 2, 1 + abs(hi) * 0.1)
        elif hi > 1e10:
            x[i] = lo + min(delta / 2, 1 + abs(lo) * 1.0)
        else:
            x[i] = (lo + hi) / 2
    return State(nvars, ncons, x, np.asarray(cvec, dtype=np.float64), A, np.asarray(b, dtype=np.float64), l, u)


def slack(l, x, u, mx):
    """slack (:137-148)"""
    d = np.minimum(mx, np.minimum(x - l, u - x))
    assert np.all(d > 0)
    return d


def max_step(l, x, u, g):
    """max-step (:120-133)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(g == 0, np.inf, np.where(g < 0, (l - x) / g, (u - x) / g))
    return s.min(initial=np.inf)


def centering_direction(l, x, u):
    """centering-direction (:150-163)"""
    out = np.where((x - l) < (u - x), np.minimum(1.0, u - x), np.maximum(-1.0, l - x))
    return np.where(np.isinf(l) & np.isinf(u), 0.0, out)


def project(scale, c, A):
    """project (:98-116): sc = scale.(-c); AD = A diag(scale); sc - AD' (AD AD')^-1 AD sc."""
    sc = scale * (-1.0 * c)
    AD2c = _mv(A, scale * sc)
    y = ns.solve_spd(ns.normal_matrix(A, scale), AD2c)
    if y is None:
        return None
    return -1.0 * (scale * _rmv(A, y)) + sc


def residual(st):
    """residual (:209-213)"""
    return st.b - _mv(st.A, st.x)


def one_repair_iteration(st, res):
    """one-repair-iteration (:226-243)"""
    sl = slack(st.l, st.x, st.u, np.sqrt(MAX_SLACK))
    t = ns.solve_spd(ns.normal_matrix(st.A, sl), res)
    dg = sl * _rmv(st.A, t)             # cholesky-ls! (:215-221)
    g = dg * sl
    step = GAMMA * min(max_step(st.l, st.x, st.u, g), 1.0 / GAMMA)
    st.x = st.x + step * g
    return True


def one_affine_scaling_iteration(st, centering=False):
    """one-affine-scaling-iteration (:165-207).  Returns `continue`."""
    x, l, u = st.x, st.l, st.u
    sl = slack(l, x, u, MAX_SLACK)
    dg = project(sl, centering_direction(l, x, u) if centering else st.c, st.A)
    if dg is None:
        return False                       # " singular " (:178-181)
    g = dg * sl
    step = GAMMA * max_step(l, x, u, g)
    norm_g, norm_dg = np.linalg.norm(g), np.linalg.norm(dg)
    descent = float(g @ st.c)
    if step > 1e10:
        raise FloatingPointError("Unbounded problem")   # (:187-188)
    if not centering:
        if norm_dg < min(1e-6, 1e-8 * len(x)) or descent > 0:
            return False
        if step * norm_g < 1e-6 or descent > 0:
            return one_affine_scaling_iteration(st, centering=True)
    st.x = x + step * g
    return True


def one_iteration(st, centering=False):
    """one-iteration (:245-263)"""
    res = residual(st)
    norm = np.linalg.norm(res)
    if norm > 1e-6 * len(res):
        st.log.append(("repair", norm))
        return one_repair_iteration(st, res)
    st.log.append(("recenter" if centering else "optimize", float(st.x @ st.c)))
    return one_affine_scaling_iteration(st, centering=centering)


def affine_scaling(st, max_iter=100000):
    """affine-scaling (:265-297): returns (c'x, x, residual, iterations)."""
    i = 0
    while i < max_iter:
        cont = one_iteration(st, (i + 1) % 16 == 0)
        res = residual(st)
        norm = np.linalg.norm(res)
        if not (cont or norm > 1e-6 * len(res)):
            return float(st.x @ st.c), st.x, res, i + 1
        i += 1
    return float(st.x @ st.c), st.x, residual(st), i
