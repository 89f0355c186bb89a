"""CPU oracle for the first-order solver of approx.lisp (APPROX: accelerated parallel proximal coordinate
descent with full-vector steps, on the penalised primal-dual formulation of a standard-form LP).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates approx.lisp constraint by constraint:
  make-quadratic / scale-quadratic / violation          approx.lisp:36-83
  accumulate-nu, make-approx-state                      approx.lisp:97-193
  primal-value, dual-value, complementarity-violation   approx.lisp:128-173
  make-approx                                           approx.lisp:195-299
  %value-&-gradient, value-&-gradient                   approx.lisp:301-351
  solve-coordinate, approx-descent, approx-iteration    approx.lisp:353-404
  project-gradient, dot-diff, project, approx           approx.lisp:406-459
Variables are stacked as v = [x (nvars) | y (ncons) | z (nvars) | w (nvars)].
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

NEG_INF, POS_INF = -np.inf, np.inf


@dataclass
class Linear:
    indices: np.ndarray
    coefs: np.ndarray


@dataclass
class Quadratic:
    indices: np.ndarray
    coefs: np.ndarray
    rhs: float
    beta: float
    scale: float = 1.0


@dataclass
class Complementarity:
    xi: int
    yi: int
    x0: float = 0.0
    y0: float = 0.0
    x_flipped: bool = False


def make_linear(pairs):
    """make-linear (:27-34): zero coefficients are dropped."""
    pairs = [(i, float(v)) for i, v in pairs if v != 0]
    return Linear(np.array([i for i, _ in pairs], dtype=np.int64), np.array([v for _, v in pairs], dtype=np.float64))


def make_quadratic(pairs, rhs=0.0, scale=None):
    """make-quadratic (:46-58): beta = number of non-zero coefficients (tau = n specialisation)."""
    pairs = [(i, float(v)) for i, v in pairs if v != 0]
    return Quadratic(np.array([i for i, _ in pairs], dtype=np.int64),
                     np.array([v for _, v in pairs], dtype=np.float64), float(rhs or 0.0), float(len(pairs)),
                     1.0 if scale is None else scale)


def scale_quadratic(q):
    """scale-quadratic (:70-74): scale = 1 / ||(coefs, rhs)||_2 when that norm exceeds 1e-6."""
    norm = math.sqrt(float(np.sum(q.coefs ** 2)) + q.rhs ** 2)
    if norm > 1e-6:
        q.scale = 1.0 / norm


def violation(q, x):
    """violation (:76-83): (a x - b) * scale."""
    return (float(q.coefs @ x[q.indices]) - q.rhs) * q.scale


@dataclass
class ApproxState:
    orig_vars: int
    orig_cons: int
    nvars: int
    constraints: list
    nu: np.ndarray
    l: np.ndarray
    u: np.ndarray
    c: list
    b: np.ndarray
    z0: float = 0.0
    log: list = field(default_factory=list)


def make_approx_state(orig_vars, orig_cons, constraints, nvars, l, u, c, b, z0=0.0):
    """make-approx-state (:175-193) with accumulate-nu (:97-113)."""
    nu = np.zeros(nvars)
    for con in constraints:
        if isinstance(con, Quadratic):
            np.add.at(nu, con.indices, con.beta * (con.coefs * con.scale) ** 2)
    return ApproxState(orig_vars, orig_cons, nvars, list(constraints), nu, l, u, c, b, z0)


def make_approx(nvars, ncons, c_pairs, triplets, b, types, l_sf, u_sf, complementarity=False, scale=True,
                l1_penalty=0.0):
    """make-approx (:195-299).  `triplets` = iterable of (row, col, value); `types` per row in
    {None, '<', '>'}; c_pairs = [(index, value)]."""
    n = 3 * nvars + ncons
    cons = [None] * (n + 2)
    for k in range(n + 2):
        cons[k] = []
    comp = {}
    l = np.full(n, NEG_INF)
    u = np.full(n, POS_INF)
    l[:nvars] = l_sf
    u[:nvars] = u_sf
    for i in range(nvars):
        yi, zi, wi = ncons + i, nvars + ncons + i, nvars + ncons + nvars + i
        li, ui = l_sf[i], u_sf[i]
        if li < -1e8:
            l[zi] = u[zi] = 0.0
        else:
            l[zi] = 0.0
            cons[yi].append((zi, 1.0))
            cons[n].append((zi, -li))
            if complementarity:
                comp[zi] = Complementarity(xi=i, yi=zi, x0=li)
        if ui > 1e8:
            l[wi] = u[wi] = 0.0
        else:
            l[wi] = 0.0
            cons[yi].append((wi, -1.0))
            cons[n].append((wi, ui))
            if complementarity:
                comp[wi] = Complementarity(xi=i, yi=wi, x0=ui, x_flipped=True)
    for (row, col, v) in triplets:
        cons[row].append((col, v))
        cons[ncons + col].append((nvars + row, v))
    out = [None] * (n + 2)
    for i in range(ncons):
        if cons[i]:
            cons[n].append((i + nvars, -b[i]))
            out[i] = make_quadratic(cons[i], rhs=b[i])
            if types[i] == "<":
                u[nvars + i] = 0.0
            elif types[i] == ">":
                l[nvars + i] = 0.0
    for i in range(ncons, ncons + nvars):
        out[i] = make_quadratic(cons[i])
    for xi, v in c_pairs:
        out[xi + ncons].rhs = float(v)
        cons[n].append((xi, v))
    for k, cp in comp.items():
        out[k] = cp
    out[n] = make_quadratic(cons[n])
    lin = []
    for i in range(nvars):
        if l[i] == NEG_INF and u[i] < POS_INF:
            lin.append((i, -l1_penalty))
        elif l[i] > NEG_INF and u[i] == POS_INF:
            lin.append((i, l1_penalty))
        else:
            lin.append((i, 0.0))
        lin.append((i + nvars + ncons, l1_penalty))
        lin.append((i + nvars + ncons + nvars, l1_penalty))
    out[n + 1] = make_linear(lin)
    if scale:
        for con in out[: n + 1]:
            if isinstance(con, Quadratic):
                scale_quadratic(con)
    constraints = [con for con in out if con is not None]
    return make_approx_state(nvars, ncons, constraints, n, l, u, list(c_pairs), np.asarray(b, dtype=np.float64))


def primal_value(st, x):
    """primal-value (:128-133)."""
    return sum(x[i] * v for i, v in st.c)


def dual_value(st, x):
    """dual-value (:135-155): b.y + sum l z [z > 0] - sum u w [w > 0]."""
    nv, nc = st.orig_vars, st.orig_cons
    acc = float(st.b @ x[nv: nv + nc])
    z = x[nv + nc: nv + nc + nv]
    w = x[nv + nc + nv: nv + nc + 2 * nv]
    lo, hi = st.l[:nv], st.u[:nv]
    zm, wm = z > 0, w > 0
    assert np.all(lo[zm] > NEG_INF) and np.all(hi[wm] < POS_INF)
    acc += float(np.sum(lo[zm] * z[zm]))
    acc -= float(np.sum(hi[wm] * w[wm]))
    return acc


def complementarity_violation(st, x):
    """complementarity-violation (:157-173)."""
    nv, nc = st.orig_vars, st.orig_cons
    z = x[nv + nc: nv + nc + nv]
    w = x[nv + nc + nv: nv + nc + 2 * nv]
    xi = x[:nv]
    lo, hi = st.l[:nv], st.u[:nv]
    with np.errstate(invalid="ignore"):
        a = np.where(z > 0, (xi - lo) * z, 0.0)
        bb = np.where(w > 0, (hi - xi) * w, 0.0)
    return float(np.sum(a) + np.sum(bb))


def _value_and_gradient_one(con, x, g):
    """%value-&-gradient (:301-336)."""
    if isinstance(con, Linear):
        if g is not None:
            np.add.at(g, con.indices, con.coefs)
        return float(con.coefs @ x[con.indices])
    if isinstance(con, Complementarity):
        xk = x[con.xi] - con.x0
        yk = x[con.yi] - con.y0
        if con.x_flipped:
            xk = -xk
        xk = max(xk, 0.0)
        yk = max(yk, 0.0)
        if g is not None:
            g[con.xi] += -yk if con.x_flipped else yk
            g[con.yi] += xk
        return yk * xk
    v = violation(con, x)
    if g is not None:
        np.add.at(g, con.indices, con.coefs * con.scale * v)
    return 0.5 * v * v


def value_and_gradient(st, x):
    """value-&-gradient (:338-351): (sum of values, gradient, max |value|)."""
    g = np.zeros(st.nvars)
    z = 0.0
    mx = 0.0
    for con in st.constraints:
        v = _value_and_gradient_one(con, x, g)
        z += v
        mx = max(mx, abs(v))
    return z, g, mx


def solve_coordinate(z, nu, theta, g, l, u):
    """solve-coordinate (:353-369), vectorised: argmin g x + (theta nu / 2)(x - z)^2 on [l, u]."""
    step = theta * nu
    with np.errstate(divide="ignore", invalid="ignore"):
        best = np.clip(z - g / step, l, u)
    zero = np.where(g < 0, u, np.where(g > 0, l, z))
    return np.where(step == 0, zero, best)


def approx_iteration(st, theta, x, z):
    """approx-iteration (:384-398) with approx-descent (:371-382)."""
    y = (1.0 - theta) * x + theta * z
    _, g, _ = value_and_gradient(st, y)
    zp = solve_coordinate(z, st.nu, theta, g, st.l, st.u)
    xn = y + theta * (zp - z)
    thn = 0.5 * (math.sqrt(theta ** 4 + 4 * theta ** 2) - theta ** 2)
    return xn, zp, thn, g


def project_gradient(st, x, g):
    """project-gradient (:400-410): x - clamp(x - g)."""
    return x - np.clip(x - g, st.l, st.u)


def approx(st, n, x=None):
    """approx (:425-459).  Returns (z, iterations done, restarts)."""
    x = np.clip(np.zeros(st.nvars) if x is None else np.asarray(x, dtype=np.float64), st.l, st.u)
    z = x.copy()
    theta = 1.0
    restarts = 0
    for i in range(n):
        x, zp, theta, _ = approx_iteration(st, theta, x, z)
        value, g, mx = value_and_gradient(st, zp)
        if float(g @ (zp - z)) > 0:
            restarts += 1
            x = z
            theta = 1.0
        else:
            z = zp
        pg = float(np.linalg.norm(project_gradient(st, z, g)))
        done = pg < 1e-10
        if i == 0 or i == n - 1 or (i + 1) % 1000 == 0 or done:
            st.log.append((i + 1, float(np.linalg.norm(g)), pg, mx, value + st.z0, complementarity_violation(st, z)))
        if done:
            return z, i + 1, restarts
    return z, n, restarts
