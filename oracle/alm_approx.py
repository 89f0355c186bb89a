"""CPU oracle for the second-generation first-order solver (alm-approx.lisp): an augmented-Lagrangian outer
loop around the APPROX inner solver.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
  make-quadratic, scale-quadratic, violation, accumulate-nu     alm-approx.lisp:36-108
  make-approx-state, dual-value, primal-value                   alm-approx.lisp:110-148
  %value-&-gradient, value-&-gradient                           alm-approx.lisp:149-196
  solve-coordinate (0.95 damping), approx-iteration             alm-approx.lisp:198-266
  project-gradient, dot-diff, project, approx                   alm-approx.lisp:268-346
  make-alm-subproblem, total-violation                          alm-approx.lisp:355-412
  make-alm, clamp, alm-iteration2, alm                          alm-approx.lisp:414-446, 492-561
The subproblem of an outer iteration is  min  c.x + lambda.(Ax - b) + (mu / 2) |Ax - b|^2,  l <= x <= u.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


@dataclass
class Subproblem:
    """approx-state of make-alm-subproblem: quadratics = rows of A with scale sqrt(weight), one linear."""
    A: sp.csr_matrix          # rows that have at least one coefficient (zeros dropped), in row order
    rhs: np.ndarray
    scale: float
    lin: np.ndarray           # c + A' lambda (dense)
    nu: np.ndarray
    l: np.ndarray
    u: np.ndarray
    z0: float
    weight: float
    lam: np.ndarray
    log: list = field(default_factory=list)


def make_alm_subproblem(A, b, c, l, u, lam, weight):
    """make-alm-subproblem (:355-403).  A: scipy sparse (ncons x nvars)."""
    A = sp.csr_matrix(A)
    A.eliminate_zeros()
    lam = np.asarray(lam, dtype=np.float64)
    cc = np.asarray(c, dtype=np.float64) + A.T @ lam
    z0 = float(lam @ b)
    scale = math.sqrt(weight)
    beta = np.diff(A.indptr).astype(np.float64)
    A2 = A.copy()
    A2.data = A2.data ** 2
    nu = (scale ** 2) * (A2.T @ beta)                      # accumulate-nu (:92-108)
    return Subproblem(A, np.asarray(b, dtype=np.float64), scale, cc, np.asarray(nu).ravel(),
                      np.asarray(l, dtype=np.float64), np.asarray(u, dtype=np.float64), -z0, weight, lam)


def value_and_gradient(sub, x):
    """value-&-gradient (:178-196): the max ignores the linear constraint."""
    viol = (sub.A @ x - sub.rhs) * sub.scale
    g = sub.A.T @ (sub.scale * viol) + sub.lin
    lin_val = float(sub.lin @ x)
    vals = 0.5 * viol * viol
    return float(vals.sum()) + lin_val, g, float(vals.max(initial=0.0))


def dual_value(sub, x):
    """dual-value (:136-140): z0 + value of the linear constraint."""
    return sub.z0 + float(sub.lin @ x)


def unscaled_violation(sub, x):
    """(violation c x nil) for every quadratic (:78-88)."""
    return sub.A @ x - sub.rhs


def solve_coordinate(z, nu, theta, g, l, u):
    """solve-coordinate (:198-216): step damped by 0.95."""
    step = theta * nu
    with np.errstate(divide="ignore", invalid="ignore"):
        best = np.clip(z - 0.95 * (g / step), l, u)
    zero = np.where(g < 0, u, np.where(g == 0, z, l))
    return np.where(step == 0, zero, best)


def approx(sub, n, x=None, accuracy=1e-5):
    """approx (:307-346): stops when i > 10 and the projected-gradient norm is below `accuracy`, or at
    the n-th iteration.  Returns (z, pg, iterations, restarts)."""
    x = np.clip(np.zeros(len(sub.l)) if x is None else np.asarray(x, dtype=np.float64), sub.l, sub.u)
    z = x.copy()
    theta = 1.0
    restarts = 0
    for i in range(n):
        y = (1.0 - theta) * x + theta * z
        _, g, _ = value_and_gradient(sub, y)
        zp = solve_coordinate(z, sub.nu, theta, g, sub.l, sub.u)
        x = y + theta * (zp - z)
        sq = theta * theta
        theta = 0.5 * (math.sqrt((4 + sq) * sq) - theta ** 2)
        value, g, mx = value_and_gradient(sub, zp)
        if float(g @ (zp - z)) > 0:
            restarts += 1
            x = z
            theta = 1.0
        else:
            z = zp
        pg = float(np.linalg.norm(z - np.clip(z - g, sub.l, sub.u)))
        done = (i > 10 and pg < accuracy) or i == n - 1
        if done:
            sub.log.append((i + 1, float(np.linalg.norm(g)), pg, mx, value + sub.z0, dual_value(sub, zp)))
            return z, pg, i + 1, restarts
    raise AssertionError("approx: n must be positive")


@dataclass
class AlmState:
    A: sp.csr_matrix
    b: np.ndarray
    c: np.ndarray
    l: np.ndarray
    u: np.ndarray
    mu: float
    omega: float
    nu: float
    multipliers: np.ndarray
    multipliers_l: np.ndarray
    multipliers_u: np.ndarray
    log: list = field(default_factory=list)


def make_alm(A, b, c, l, u, types, mu=10.0, multipliers=None):
    """make-alm (:424-446)."""
    m = len(b)
    low = np.full(m, -np.inf)
    high = np.full(m, np.inf)
    for i, t in enumerate(types):
        if t == "<":
            low[i] = 0.0
        elif t == ">":
            high[i] = 0.0
    return AlmState(sp.csr_matrix(A), np.asarray(b, dtype=np.float64), np.asarray(c, dtype=np.float64),
                    np.asarray(l, dtype=np.float64), np.asarray(u, dtype=np.float64), float(mu), 1.0 / mu,
                    (1.0 / mu) ** 0.1, np.zeros(m) if multipliers is None else np.asarray(multipliers, float),
                    low, high)


def alm_iteration2(st, x, precision=None, max_inner=1000000):
    """alm-iteration2 (:492-537).  Returns (x, violation vector, pg, dual value, inner iterations)."""
    sub = make_alm_subproblem(st.A, st.b, st.c, st.l, st.u, st.multipliers, st.mu)
    violation0 = None if x is None else float(np.linalg.norm(unscaled_violation(sub, x)))
    x, pg, inner, _ = approx(sub, max_inner, x, precision if precision is not None else max(st.omega, 1e-6))
    value = dual_value(sub, x)
    violation = unscaled_violation(sub, x)
    improvement = None if not violation0 else float(np.linalg.norm(violation)) / violation0
    st.multipliers = np.maximum(st.multipliers_l, np.minimum(st.multipliers + st.mu * violation, st.multipliers_u))
    factor = max(1.0, min(2.0 * improvement, 2.0)) if improvement else 1.0
    st.mu = min(st.mu * factor, 1e7)
    st.nu = 1.0 / st.mu ** 0.1
    st.omega = max(1.0 / st.mu, 1e-6)
    st.log.append((float(np.abs(violation).max(initial=0.0)), float(np.linalg.norm(violation)), pg, value, st.mu, inner))
    return x, violation, pg, value, inner


def alm(st, x0=None, maxiter=None, max_inner=1000000):
    """alm (:539-561).  Returns (outer iterations, inner iterations, |violation|_inf, pg, dual value, x)."""
    x, v, pg, z = x0, None, None, None
    accuracy = math.inf
    total_inner = 0
    i = 0
    for i in range(maxiter or 10000):
        x, vv, pg, z, inner = alm_iteration2(st, x, min(accuracy, st.omega), max_inner)
        total_inner += inner
        v = float(np.abs(vv).max(initial=0.0))
        accuracy = min(accuracy, max(1e-5, v))
        if v < 1e-5:
            accuracy = 1e-5
        if not (v > 1e-5 or pg > 1e-5):
            return i, total_inner, v, pg, z, x
    return i + 1, total_inner, v, pg, z, x


# ---- the other outer-loop variants of alm-approx.lisp ----------------------------------------------
def alm_iteration(st, x, precision=None, max_inner=1000000):
    """alm-iteration (:448-490): "minor" step (multipliers only) when the violation is below nu, else "major"
    (mu grows by 1.5).  Returns (x, violation, dual value, kind)."""
    sub = make_alm_subproblem(st.A, st.b, st.c, st.l, st.u, st.multipliers, st.mu)
    x, pg, inner, _ = approx(sub, max_inner, x, precision if precision is not None else max(st.omega, 1e-5))
    value = dual_value(sub, x)
    violation = unscaled_violation(sub, x)
    vnorm = float(np.linalg.norm(violation))
    st.multipliers = st.multipliers + st.mu * violation
    if vnorm < st.nu:
        st.nu = st.nu / st.mu ** 0.9
        st.omega = max(st.omega / st.mu, 1e-5)
        kind = "minor"
    else:
        st.mu = min(1.5 * st.mu, 1e6)
        st.nu = 1.0 / st.mu ** 0.1
        st.omega = max(1.0 / st.mu, 1e-5)
        kind = "major"
    st.log.append((vnorm, pg, value, kind, inner))
    return x, violation, value, kind


def next_extrapolation(weight):
    """next-extrapolation (:563-564)."""
    return 0.5 * (1 + math.sqrt(1 + 4 * weight * weight))


def extrapolate(weight, prev, accelerated, current):
    """extrapolate (:566-577)."""
    nxt = next_extrapolation(weight)
    return current + ((weight - 1) / nxt) * (current - prev) + (weight / nxt) * (current - accelerated)


def aalm(st, x0=None, maxiter=None, max_inner=1000000):
    """aalm (:579-610), the accelerated outer loop ("not very good" in the source)."""
    x, v, pg, z = x0, None, None, None
    accuracy = math.inf
    prev_multipliers = st.multipliers
    extrapolation = 1.0
    total = 0
    i = 0
    for i in range(maxiter or 10000):
        prev_accelerated = st.multipliers
        if i > 0:
            extrapolation = next_extrapolation(extrapolation)
        x, vv, pg, z, inner = alm_iteration2(st, x, min(accuracy, st.omega), max_inner)
        total += inner
        v = float(np.abs(vv).max(initial=0.0))
        accuracy = min(accuracy, max(1e-6, v))
        if v < 1e-5:
            accuracy = 1e-6
        prev_multipliers, st.multipliers = st.multipliers, extrapolate(extrapolation, prev_multipliers,
                                                                       prev_accelerated, st.multipliers)
        if not (v > 1e-5 or (pg > 1e-5 and pg > 2e-6 * (1 + abs(z)))):
            return i, total, v, pg, z, x
    return i + 1, total, v, pg, z, x


def adcd_iteration(st, x):
    """adcd-iteration (:612-656): short inner solves (100 iterations, 10000 once close) to accuracy 1e-2.
    Returns (x, violation, done)."""
    sub = make_alm_subproblem(st.A, st.b, st.c, st.l, st.u, st.multipliers, st.mu)
    close = x is not None and float(np.linalg.norm(unscaled_violation(sub, x))) < 5e-2
    x, pg, inner, _ = approx(sub, 10000 if close else 100, x, 1e-2)
    violation = unscaled_violation(sub, x)
    vnorm = float(np.linalg.norm(violation))
    out_close = pg < 5e-2
    almost = vnorm < 5e-2
    if pg < 1e-2 and vnorm < 1e-2:
        return x, violation, True
    st.multipliers = st.multipliers + ((1.0 if out_close else 0.5) * st.mu) * violation
    st.mu = min(1e6, (1.0 if (out_close and almost) else (10.0 if out_close else 1.0)) * st.mu)
    st.nu = 1.0 / st.mu ** 0.1
    st.omega = 1.0 / st.mu
    return x, violation, False
