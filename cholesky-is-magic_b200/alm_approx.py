"""Host-side mirror of alm-approx.lisp: the augmented-Lagrangian outer loop (`make-alm`, `alm-iteration2`,
`alm`, alm-approx.lisp:414-446, 492-561) around the APPROX inner solver, which runs on the GPU
(nes_approx_* with variant 1).  The subproblems of all outer iterations share one device-resident
constraint matrix; an outer iteration only sends the new linear term and reads the violation back."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import nes
from .sparse_cholesky import cholmod_common


class AlmState:
    """alm-state (alm-approx.lisp:414-422) + the device handle of the subproblem."""

    def __init__(self, sf, mu=10.0, multipliers=None):
        com = cholmod_common()
        self.sf = sf
        m, n = sf.ncons, sf.nvars
        self.mu = float(mu)
        self.omega = 1.0 / mu
        self.nu = (1.0 / mu) ** 0.1
        self.multipliers = np.zeros(m) if multipliers is None else np.asarray(multipliers, dtype=np.float64).copy()
        self.multipliers_l = np.full(m, -np.inf)
        self.multipliers_u = np.full(m, np.inf)
        for i, t in enumerate(sf.type if sf.type is not None else [None] * m):
            if t == "<":
                self.multipliers_l[i] = 0.0
            elif t == ">":
                self.multipliers_u[i] = 0.0
        if sf.A is not None:
            r, c_, v = np.asarray(sf.A.row), np.asarray(sf.A.col), np.asarray(sf.A.value, dtype=np.float64)
        else:
            r, c_ = np.nonzero(sf.A_dense)
            v = sf.A_dense[r, c_]
        keep = v != 0
        self.A = nes.Matrix.from_triplets(com, r[keep].astype(np.int32), c_[keep].astype(np.int32), v[keep], m, n)
        self.b = np.asarray(sf.b, dtype=np.float64)
        self.c = np.asarray(sf.c_dense(), dtype=np.float64)
        l = np.ascontiguousarray(sf.l, dtype=np.float64)
        u = np.ascontiguousarray(sf.u, dtype=np.float64)
        lin0 = np.zeros(n)
        self.ptr = com.lib.nes_approx_create(self.A.ptr, self.b.ctypes.data_as(nes._dp), lin0.ctypes.data_as(nes._dp),
                                             l.ctypes.data_as(nes._dp), u.ctypes.data_as(nes._dp), None, None, None,
                                             None, 0, None, 0.0, 0, 0.0, com.ptr)
        if not self.ptr:
            self.A.free()
            raise nes.NesError(f"nes_approx_create failed: {com.error()}")
        self.log = []

    def free(self):
        com = cholmod_common()
        if self.ptr:
            h = C.c_void_p(self.ptr)
            assert com.lib.nes_approx_free(C.byref(h), com.ptr) != 0
            self.ptr = None
        if self.A is not None:
            self.A.free()
            self.A = None


def make_alm(sf, mu=10.0, multipliers=None):
    """make-alm (alm-approx.lisp:424-446)."""
    return AlmState(sf, mu, multipliers)


def alm_iteration2(st, x, precision=None, max_inner=1000000):
    """alm-iteration2 (alm-approx.lisp:492-537): (x, violation, pg, dual value, inner iterations)."""
    com = cholmod_common()
    lin = st.c + st.A.sdmult(st.multipliers, transpose=True)                 # c + A' lambda
    z0 = -float(st.multipliers @ st.b)
    com.check(com.lib.nes_approx_set_subproblem(st.ptr, math.sqrt(st.mu), lin.ctypes.data_as(nes._dp), z0, com.ptr),
              "nes_approx_set_subproblem")
    violation0 = None if x is None else float(np.linalg.norm(st.A.sdmult(x) - st.b))
    acc = precision if precision is not None else max(st.omega, 1e-6)
    com.check(com.lib.nes_approx_set_variant(st.ptr, 1, acc, com.ptr), "nes_approx_set_variant")
    z = np.empty(st.sf.nvars)
    it, rs = C.c_int(), C.c_int()
    stats = np.zeros(7)
    x0 = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    com.check(com.lib.nes_approx_solve(st.ptr, max_inner, None if x0 is None else x0.ctypes.data_as(nes._dp),
                                       z.ctypes.data_as(nes._dp), C.byref(it), C.byref(rs),
                                       stats.ctypes.data_as(nes._dp), com.ptr), "nes_approx_solve")
    x, pg = z, float(stats[1])
    value = z0 + float(lin @ x)                                              # dual-value (:136-140)
    violation = st.A.sdmult(x) - st.b
    improvement = None if not violation0 else float(np.linalg.norm(violation)) / violation0
    st.multipliers = np.maximum(st.multipliers_l, np.minimum(st.multipliers + st.mu * violation, st.multipliers_u))
    st.mu = min(st.mu * (max(1.0, min(2.0 * improvement, 2.0)) if improvement else 1.0), 1e7)
    st.nu = 1.0 / st.mu ** 0.1
    st.omega = max(1.0 / st.mu, 1e-6)
    st.log.append((float(np.abs(violation).max(initial=0.0)), float(np.linalg.norm(violation)), pg, value, st.mu,
                   it.value))
    return x, violation, pg, value, it.value


def alm(st, x0=None, maxiter=None, max_inner=1000000):
    """alm (alm-approx.lisp:539-561): (outer iterations, inner iterations, |violation|_inf, pg, value, x)."""
    x, v, pg, z = x0, None, None, None
    accuracy = math.inf
    total = 0
    i = 0
    for i in range(maxiter or 10000):
        x, vv, pg, z, inner = alm_iteration2(st, x, min(accuracy, st.omega), max_inner)
        total += inner
        v = float(np.abs(vv).max(initial=0.0))
        accuracy = min(accuracy, max(1e-5, v))
        if v < 1e-5:
            accuracy = 1e-5
        if not (v > 1e-5 or pg > 1e-5):
            return i, total, v, pg, z, x
    return i + 1, total, v, pg, z, x


# ---- the other outer-loop variants of alm-approx.lisp (host loops over the same device solver) -------
def _inner(st, x, accuracy, max_inner):
    """make-alm-subproblem + approx on the device; returns (x, pg, dual value, violation, inner iterations)."""
    com = cholmod_common()
    lin = st.c + st.A.sdmult(st.multipliers, transpose=True)
    z0 = -float(st.multipliers @ st.b)
    com.check(com.lib.nes_approx_set_subproblem(st.ptr, math.sqrt(st.mu), lin.ctypes.data_as(nes._dp), z0, com.ptr),
              "nes_approx_set_subproblem")
    com.check(com.lib.nes_approx_set_variant(st.ptr, 1, accuracy, com.ptr), "nes_approx_set_variant")
    z = np.empty(st.sf.nvars)
    it, rs = C.c_int(), C.c_int()
    stats = np.zeros(7)
    x0 = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    com.check(com.lib.nes_approx_solve(st.ptr, max_inner, None if x0 is None else x0.ctypes.data_as(nes._dp),
                                       z.ctypes.data_as(nes._dp), C.byref(it), C.byref(rs),
                                       stats.ctypes.data_as(nes._dp), com.ptr), "nes_approx_solve")
    return z, float(stats[1]), z0 + float(lin @ z), st.A.sdmult(z) - st.b, it.value


def alm_iteration(st, x, precision=None, max_inner=1000000):
    """alm-iteration (alm-approx.lisp:448-490): (x, violation, dual value, "minor" | "major")."""
    x, pg, value, violation, inner = _inner(st, x, precision if precision is not None else max(st.omega, 1e-5),
                                            max_inner)
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
    """next-extrapolation (alm-approx.lisp:563-564)."""
    return 0.5 * (1 + math.sqrt(1 + 4 * weight * weight))


def extrapolate(weight, prev, accelerated, current):
    """extrapolate (alm-approx.lisp:566-577)."""
    nxt = next_extrapolation(weight)
    return current + ((weight - 1) / nxt) * (current - prev) + (weight / nxt) * (current - accelerated)


def aalm(st, x0=None, maxiter=None, max_inner=1000000):
    """aalm (alm-approx.lisp:579-610)."""
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
    """adcd-iteration (alm-approx.lisp:612-656): (x, violation, done)."""
    close = x is not None and float(np.linalg.norm(st.A.sdmult(x) - st.b)) < 5e-2
    x, pg, value, violation, inner = _inner(st, x, 1e-2, 10000 if close else 100)
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
