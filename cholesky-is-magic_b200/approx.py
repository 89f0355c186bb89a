"""Host-side mirror of approx.lisp (the first-order APPROX solver) over the C ABI: `make_approx` lays the
penalised primal-dual formulation of a standard-form LP out as arrays (index bookkeeping only -- the
arithmetic, including scale-quadratic and accumulate-nu, happens on the GPU), `approx` runs the solver
(nes_approx_solve).  Same names and argument meaning as approx.lisp:195-299, 425-459."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import nes
from .sparse_cholesky import cholmod_common


class ApproxState:
    """approx-state (approx.lisp:115-126) with its device handle."""

    def __init__(self, sf, K, handle, n, l, u, rhs, lin, comps):
        self.sf, self.K, self.ptr, self.nvars = sf, K, handle, n
        self.orig_vars, self.orig_cons = sf.nvars, sf.ncons
        self.l, self.u, self.rhs, self.lin, self.comps = l, u, rhs, lin, comps
        self.log = []

    def _get(self, which, count):
        com = cholmod_common()
        out = np.empty(count)
        com.check(com.lib.nes_approx_get(self.ptr, ord(which), out.ctypes.data_as(nes._dp), com.ptr), "nes_approx_get")
        return out

    @property
    def nu(self):
        return self._get("n", self.nvars)

    @property
    def scale(self):
        """Scales of the quadratics in constraint order (the gap row, kept dense, is last)."""
        return np.concatenate([self._get("s", len(self.rhs)), self._get("d", 2)[:1]])

    def free(self):
        com = cholmod_common()
        if self.ptr:
            h = C.c_void_p(self.ptr)
            assert com.lib.nes_approx_free(C.byref(h), com.ptr) != 0
            self.ptr = None
        if self.K is not None:
            self.K.free()
            self.K = None


def make_approx(sf, complementarity=False, scale=True, l1_penalty=0.0):
    """make-approx (approx.lisp:195-299): variables [x | y | z | w]; quadratics |Ax - b|, |A'y + z - w - c| and
    the duality-gap row; optional complementarity constraints z (x - l), w (u - x)."""
    com = cholmod_common()
    nvars, ncons = sf.nvars, sf.ncons
    n = 3 * nvars + ncons
    l = np.full(n, -np.inf)
    u = np.full(n, np.inf)
    l[:nvars] = sf.l
    u[:nvars] = sf.u
    rows, cols, vals = [], [], []            # triplets of K in constraint-array positions (compacted below)
    gap = n                                   # position of the duality-gap quadratic
    comp = []
    for i in range(nvars):
        yi, zi, wi = ncons + i, nvars + ncons + i, nvars + ncons + nvars + i
        li, ui = sf.l[i], sf.u[i]
        if li < -1e8:
            l[zi] = u[zi] = 0.0
        else:
            l[zi] = 0.0
            rows += [yi, gap]; cols += [zi, zi]; vals += [1.0, -li]
            if complementarity:
                comp.append((i, zi, li, 0))
        if ui > 1e8:
            l[wi] = u[wi] = 0.0
        else:
            l[wi] = 0.0
            rows += [yi, gap]; cols += [wi, wi]; vals += [-1.0, ui]
            if complementarity:
                comp.append((i, wi, ui, 1))
    if sf.A is not None:
        ar, ac, av = np.asarray(sf.A.row), np.asarray(sf.A.col), np.asarray(sf.A.value, dtype=np.float64)
    else:
        ar, ac = np.nonzero(sf.A_dense)
        av = sf.A_dense[ar, ac]
    rows += ar.tolist() + (ncons + ac).tolist()
    cols += ac.tolist() + (nvars + ar).tolist()
    vals += av.tolist() + av.tolist()
    rhs = np.zeros(n + 1)
    has_pairs = np.zeros(ncons, dtype=bool)
    has_pairs[ar[av != 0]] = True
    types = list(sf.type) if sf.type is not None else [None] * ncons
    for i in range(ncons):
        if has_pairs[i]:
            rows.append(gap); cols.append(i + nvars); vals.append(-sf.b[i])
            rhs[i] = sf.b[i]
            if types[i] == "<":
                u[nvars + i] = 0.0
            elif types[i] == ">":
                l[nvars + i] = 0.0
    for xi, v in sf.c:
        rhs[xi + ncons] = v
        rows.append(gap); cols.append(xi); vals.append(v)
    lin = np.zeros(n)
    for i in range(nvars):
        if l[i] == -np.inf and u[i] < np.inf:
            lin[i] = -l1_penalty
        elif l[i] > -np.inf and u[i] == np.inf:
            lin[i] = l1_penalty
        lin[i + nvars + ncons] = l1_penalty
        lin[i + nvars + ncons + nvars] = l1_penalty
    rows, cols, vals = np.asarray(rows), np.asarray(cols), np.asarray(vals, dtype=np.float64)
    keep = vals != 0                           # make-quadratic drops zero coefficients (:47)
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    # the duality-gap quadratic touches every variable: it travels as a dense coefficient vector
    isgap = rows == gap
    gap_row = np.zeros(n)
    np.add.at(gap_row, cols[isgap], vals[isgap])
    rows, cols, vals = rows[~isgap], cols[~isgap], vals[~isgap]
    # remaining quadratics in constraint order: non-empty primal rows, then all dual rows
    present = np.zeros(n, dtype=bool)
    present[:ncons] = has_pairs
    present[ncons: ncons + nvars] = True
    newrow = np.cumsum(present) - 1
    R = int(present.sum())
    K = nes.Matrix.from_triplets(com, newrow[rows].astype(np.int32), cols.astype(np.int32), vals, R, n)
    rhs_c = np.ascontiguousarray(rhs[:n][present])
    cx = np.array([c_[0] for c_ in comp], dtype=np.int32)
    cy = np.array([c_[1] for c_ in comp], dtype=np.int32)
    cx0 = np.array([c_[2] for c_ in comp], dtype=np.float64)
    cf = np.array([c_[3] for c_ in comp], dtype=np.int32)
    ptr = lambda a, t: a.ctypes.data_as(t) if len(a) else None
    h = com.lib.nes_approx_create(K.ptr, rhs_c.ctypes.data_as(nes._dp), lin.ctypes.data_as(nes._dp),
                                  l.ctypes.data_as(nes._dp), u.ctypes.data_as(nes._dp), ptr(cx, nes._ip),
                                  ptr(cy, nes._ip), ptr(cx0, nes._dp), ptr(cf, nes._ip), len(comp),
                                  gap_row.ctypes.data_as(nes._dp), 0.0, 1 if scale else 0, 0.0, com.ptr)
    if not h:
        K.free()
        raise nes.NesError(f"nes_approx_create failed: {com.error()}")
    return ApproxState(sf, K, h, n, l, u, rhs_c, lin, comp)


def value_and_gradient(state, x):
    """value-&-gradient (approx.lisp:338-351): (value, gradient, max |constraint value|)."""
    com = cholmod_common()
    x = np.ascontiguousarray(x, dtype=np.float64)
    g = np.empty(state.nvars)
    val, mx = C.c_double(), C.c_double()
    com.check(com.lib.nes_approx_value_gradient(state.ptr, x.ctypes.data_as(nes._dp), C.byref(val),
                                                g.ctypes.data_as(nes._dp), C.byref(mx), com.ptr),
              "nes_approx_value_gradient")
    return val.value, g, mx.value


def approx(state, n, x=None):
    """approx (approx.lisp:425-459).  Returns (z, iterations, restarts, stats) with stats =
    (|g|, projected-gradient norm, max constraint value, value + z0, last g.(zp - z), theta)."""
    com = cholmod_common()
    z = np.empty(state.nvars)
    it, rs = C.c_int(), C.c_int()
    stats = np.zeros(7)
    x0 = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    com.check(com.lib.nes_approx_solve(state.ptr, n, None if x0 is None else x0.ctypes.data_as(nes._dp),
                                       z.ctypes.data_as(nes._dp), C.byref(it), C.byref(rs),
                                       stats.ctypes.data_as(nes._dp), com.ptr), "nes_approx_solve")
    return z, it.value, rs.value, stats


def primal_value(state, x):
    """primal-value (approx.lisp:128-133)."""
    return float(sum(x[i] * v for i, v in state.sf.c))


def dual_value(state, x):
    """dual-value (approx.lisp:135-155)."""
    nv, nc = state.orig_vars, state.orig_cons
    z, w = x[nv + nc: 2 * nv + nc], x[2 * nv + nc: 3 * nv + nc]
    lo, hi = state.l[:nv], state.u[:nv]
    zm, wm = z > 0, w > 0
    return float(state.sf.b @ x[nv: nv + nc] + np.sum(lo[zm] * z[zm]) - np.sum(hi[wm] * w[wm]))
