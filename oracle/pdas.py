"""Oracle (test infrastructure): primal-dual affine scaling, restating
primal-dual-affine-scaling.lisp function by function in NumPy (see oracle/__init__.py: parity is
unpinned by the reference; this file transcribes its literal rules).

A is a dense ndarray or a scipy.sparse matrix.  CHOLMOD's factorize/solve is played by OpenBLAS
dpotrf/dpotrs on the explicitly formed normal matrix (oracle/newton_solve.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import newton_solve as ns

CLAMP = 1e8  # *clamp*, primal-dual-affine-scaling.lisp:37


def clamp(v):
    """clamp (:39-45)"""
    return np.maximum(-CLAMP, np.minimum(np.asarray(v, dtype=np.float64), CLAMP))


def scale_constraints_dense(A, b):
    """scale-constraints (:50-73) for a dense A (every row appears in a triplet)."""
    rowmax = np.abs(A).max(axis=1)
    s = np.ones(len(b))
    sel = rowmax >= 1e-6
    s[sel] = 1.0 / rowmax[sel]
    return A * s[:, None], b * s


def scale_constraints_sparse(A, b):
    A = sp.csc_matrix(A)
    rowmax = np.asarray(abs(A).max(axis=1).todense()).ravel()
    present = np.asarray((A != 0).sum(axis=1)).ravel() > 0
    s = np.ones(len(b))
    sel = present & (rowmax >= 1e-6)
    s[sel] = 1.0 / rowmax[sel]
    return sp.csc_matrix(sp.diags(s) @ A), b * np.where(present, s, 1.0)


@dataclass
class State:
    """pdas-state (:8-15)"""
    nvars: int
    ncons: int
    c: np.ndarray
    A: object
    b: np.ndarray
    l: np.ndarray
    u: np.ndarray
    x: np.ndarray
    y: np.ndarray
    w: np.ndarray
    z: np.ndarray
    filters: bool = False
    log: list = field(default_factory=list)


def make_pdas(nvars, ncons, cvec, A, b, sf_l, sf_u, filters=None, scale=True):
    """make-pdas (:75-133)."""
    cvec = np.asarray(cvec, dtype=np.float64)
    sf_l = np.asarray(sf_l, dtype=np.float64)
    sf_u = np.asarray(sf_u, dtype=np.float64)
    l, u = clamp(sf_l), clamp(sf_u)
    x = np.zeros(nvars)
    for i in range(nvars):
        if u[i] - l[i] < 1e-6:          # :88-94
            l[i] -= 5e-7
            u[i] += 5e7
        lo, hi = sf_l[i], sf_u[i]       # :95-107, unclamped bounds
        delta = hi - lo
        if lo < -1e10 and hi > 1e10:
            xi = 0.0
        elif lo < -1e6:
            xi = hi - min(delta / 2, 1 + abs(hi) * 0.1)
        elif hi > 1e6:
            xi = lo + min(delta / 2, 1 + abs(lo) * 0.1)
        else:
            xi = (lo + hi) / 2
        x[i] = xi
    z, w = np.empty(nvars), np.empty(nvars)
    for i in range(nvars):              # :108-118
        ci = cvec[i]
        if ci == 0:
            z[i], w[i] = 1.0, 1.0
        elif ci < 0:
            z[i], w[i] = 1.0, 1.0 + (-ci)
        else:
            z[i], w[i] = 1.0 + ci, 1.0
    b = np.asarray(b, dtype=np.float64)
    if scale:
        if sp.issparse(A):
            A, b = scale_constraints_sparse(A, b)
        else:
            A, b = scale_constraints_dense(np.asarray(A, dtype=np.float64), b)
    if filters is None:
        filters = bool(sp.issparse(A))
    return State(nvars, ncons, cvec, A, np.array(b, copy=True), l, u, x, np.zeros(ncons), w, z, filters)


def _mv(A, x):
    return np.asarray(A @ x).ravel()


def _rmv(A, y):
    return np.asarray(A.T @ y).ravel()


def violation(st):
    """violation (:135-150)"""
    x = st.x
    l = x - st.l
    u = st.u - x
    return (l, u, st.w * u, st.z * l, _mv(st.A, x) - st.b,
            (st.z + _rmv(st.A, st.y)) - (st.w + st.c))


def box_step(l, u, dx):
    """box-step (:166-180)"""
    mx = np.inf
    for li, di, ui in zip(l, dx, u):
        di = -di
        assert li > 0 and ui > 0
        if di == 0:
            continue
        if di < 0:
            mx = min(mx, li / (-di))
        else:
            mx = min(mx, ui / di)
    return mx


def box_step_vec(l, u, dx):
    t = -dx
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(t < 0, l / (-t), np.inf)
        b = np.where(t > 0, u / t, np.inf)
    return min(a.min(initial=np.inf), b.min(initial=np.inf))


def pos_step(v, dv):
    """pos-step (:182-192)"""
    t = -dv
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(t < 0, -(v / t), np.inf)
    return r.min(initial=np.inf)


def pdas_step(l, u, st, dw, dx, dy, dz):
    """pdas-step (:194-198)"""
    return min(box_step_vec(l, u, dx), pos_step(st.w, dw), pos_step(st.z, dz))


def apply_step(st, step, dw, dx, dy, dz):
    """apply-step (:200-207): axpy! (- step)"""
    st.w = st.w + (-step) * dw
    st.x = st.x + (-step) * dx
    st.y = st.y + (-step) * dy
    st.z = st.z + (-step) * dz
    return st


def cholesky_ls(A, s, x):
    """cholesky-ls! (:223-233): N = A diag(s); N' (N N')^-1 x"""
    M = ns.normal_matrix(A, s)
    t = ns.solve_spd(M, x)
    if t is None:
        return None
    return s * _rmv(A, t)


def slack(l, x, u, mx):
    """slack (:235-246)"""
    d = np.minimum(mx, np.minimum(x - l, u - x))
    assert np.all(d > 0)
    return d


def residual(st):
    """residual (:248-251)"""
    return st.b - _mv(st.A, st.x)


def max_step(l, x, u, g):
    """max-step (:253-266)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(g == 0, np.inf, np.where(g < 0, (l - x) / g, (u - x) / g))
    return s.min(initial=np.inf)


def one_repair_iteration(st):
    """one-repair-iteration (:268-288)"""
    x, l, u = st.x, st.l, st.u
    sl = slack(l, x, u, 1e4)
    r = residual(st)
    dg = cholesky_ls(st.A, sl, r.copy())
    g = dg * sl
    gamma = 0.9
    step = gamma * min(max_step(l, x, u, g), 1.0 / gamma)
    st.x = np.maximum(x + step * g, 1e-4)
    return np.linalg.norm(g), step


def centering_direction(l, x, u):
    """centering-direction (:290-303)"""
    out = np.where((x - l) < (u - x), np.minimum(1.0, u - x), np.maximum(-1.0, l - x))
    return np.where(np.isinf(l) & np.isinf(u), 0.0, out)


def primal_project(scale, c, A):
    """primal-project (:305-317): sc = scale.(-c); sc - A'^T (A' A'^T)^-1 A' sc, A' = A diag(scale)"""
    sc = scale * (-1.0 * c)
    Asc = _mv(A, scale * sc)
    M = ns.normal_matrix(A, scale)
    proj = ns.solve_spd(M, Asc)
    if proj is None:
        return None
    return -1.0 * (scale * _rmv(A, proj)) + sc


def one_pdas_iteration(st, repair):
    """one-pdas-iteration (:319-383).  Returns (gap, dobj, step|None)."""
    l, u, wu, zl, rp, rd = violation(st)
    assert np.all(l > 0) and np.all(u > 0)
    pobj = float(st.c @ st.x)
    dobj = float(st.b @ st.y) + float(st.l @ st.z) + (-float(st.u @ st.w))
    viol = [np.abs(rp).max(), np.abs(rd).max(), np.abs(wu).max(), np.abs(zl).max()]
    primal_feasible = viol[0] < 1e-2
    gap = abs(pobj - dobj) / max(abs(pobj), abs(dobj), 1.0)
    entry = {"pobj": pobj, "dobj": dobj, "violations": viol, "gap": gap}
    st.log.append(entry)
    if not primal_feasible:
        entry["branch"] = "repair"
        one_repair_iteration(st)
        return gap, dobj, None
    if repair:
        entry["branch"] = "recentre"
        st.w = st.w + 1e-4
        st.z = st.z + 1e-4
        sl = slack(st.l, st.x, st.u, 1e4)
        dx = primal_project(sl, centering_direction(st.l, st.x, st.u), st.A)
        dx = dx * sl
        step = 0.5 * max_step(st.l, st.x, st.u, dx)
        st.x = st.x + step * dx
        return gap, dobj, None
    res = ns.solve_kkt_newton(l, u, st.w, st.z, st.A, wu, zl, rp, rd, filters=st.filters,
                              return_intermediates=True)
    if res is None:
        raise FloatingPointError("solve-delta-y returned NIL (Cholesky failed)")
    dw, dx, dy, dz, inter = res
    step = pdas_step(l, u, st, dw, dx, dy, dz)
    entry.update(branch="newton", step=step, theta=inter["theta"], rhs=inter["rhs"], dy=dy)
    alpha = min(1.0, 0.9 * step)
    assert 0.0 < alpha <= 1.0
    apply_step(st, alpha, dw, dx, dy, dz)
    return gap, dobj, step


def pdas(st, max_iter=None):
    """pdas (:385-396).  Returns (dobj, gap, iterations)."""
    repair = False
    i = 0
    while max_iter is None or i < max_iter:
        i += 1
        viol, obj, step = one_pdas_iteration(st, repair)
        repair = step is not None and step < 1e-6
        if viol < 1e-4:
            break
    return obj, viol, i
