"""Primal-dual affine scaling driver (primal-dual-affine-scaling.lisp), host-side mirror.

    min c'x  s.t.  Ax = b,  l < x < u;   y: multipliers of Ax = b;  z, w: multipliers of the bounds.

The control flow, constants and return values follow the Lisp function by function; every vector
lives on the GPU inside a `nes_pdas` handle and each step is one C-ABI call that returns scalars.
`pdas()` can also hand the whole loop to the library (`native_loop=True` -> nes_pdas_solve, the
same loop in C++), which is what bench.py times.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import nes
from .sparse_cholesky import cholmod_common, make_sparse_from_triplet_vector
from .standard_form import StandardForm, Triplets

CLAMP = 1e8  # *clamp* (primal-dual-affine-scaling.lisp:37)


def clamp(vector):
    """clamp (:39-45)."""
    return np.maximum(-CLAMP, np.minimum(np.asarray(vector, dtype=np.float64), CLAMP))


def scale_constraints(triplets: Triplets, rhs):
    """scale-constraints (:50-73): rows scaled by 1/max|a_ij| (1 when the max is < 1e-6); only rows
    that appear in a triplet are touched.  Returns (triplets, rhs) copies."""
    rhs = np.array(rhs, dtype=np.float64, copy=True)
    rowmax = np.zeros(len(rhs))
    np.maximum.at(rowmax, triplets.row, np.abs(triplets.value))
    present = np.zeros(len(rhs), dtype=bool)
    present[triplets.row] = True
    scale = np.ones(len(rhs))
    sel = present & (rowmax >= 1e-6)
    scale[sel] = 1.0 / rowmax[sel]
    rhs *= np.where(present, scale, 1.0)
    return Triplets(triplets.row.copy(), triplets.col.copy(), triplets.value * scale[triplets.row]), rhs


@dataclass
class PdasState:
    """pdas-state (:8-15).  x, y, w, z live on the device inside `handle`; use get()/set()."""
    nvars: int
    ncons: int
    c: np.ndarray
    triplets: Triplets | None
    b: np.ndarray
    l: np.ndarray
    u: np.ndarray
    x0: np.ndarray
    y0: np.ndarray
    w0: np.ndarray
    z0: np.ndarray
    A_dense: np.ndarray | None = None
    A_generated: tuple | None = None      # (seed,) -> matrix regenerated on the device
    filters: bool | None = None
    _device_row_scale: str | None = None
    _A: nes.Matrix | None = None          # pdas-%A
    _handle: int | None = None
    log: list = field(default_factory=list)

    # -- pdas-A (:17-23): build the device matrix on first use --------------------------------
    def A(self):
        if self._A is None:
            c = cholmod_common()
            if self.A_generated is not None:
                self._A = nes.Matrix.generate_dense(c, self.ncons, self.nvars, self.A_generated[0])
            elif self.A_dense is not None:
                self._A = nes.Matrix.from_dense(c, self.A_dense)
            else:
                self._A = make_sparse_from_triplet_vector(self.ncons, self.nvars, self.triplets)
        return self._A

    def handle(self):
        if self._handle is None:
            c = cholmod_common()
            A = self.A()
            _ensure_device_row_scale(self)
            filters = (not A.is_dense) if self.filters is None else self.filters
            keep = [nes.vec(v)[0] for v in (self.c, self.b, self.l, self.u, self.x0, self.y0, self.w0, self.z0)]
            ptrs = [k.ctypes.data_as(nes._dp) for k in keep]
            h = c.lib.nes_pdas_create(A.ptr, *ptrs, 1 if filters else 0, c.ptr)
            if not h:
                raise nes.NesError(f"nes_pdas_create failed: {c.error()}")
            self._handle = h
        return self._handle

    def get(self, which):
        c = cholmod_common()
        n = self.ncons if which in ("y", "Y", "p") else self.nvars
        out = np.empty(n)
        c.check(c.lib.nes_pdas_get(self.handle(), ord(which), out.ctypes.data_as(nes._dp), c.ptr), "nes_pdas_get")
        return out

    def set(self, which, value):
        c = cholmod_common()
        v, p = nes.vec(value)
        c.check(c.lib.nes_pdas_set(self.handle(), ord(which), p, c.ptr), "nes_pdas_set")

    x = property(lambda self: self.get("x"))
    y = property(lambda self: self.get("y"))
    w = property(lambda self: self.get("w"))
    z = property(lambda self: self.get("z"))


def free_pdas_A(state: PdasState):
    """free-pdas-A (:32-35) plus the device state."""
    c = cholmod_common()
    if state._handle is not None:
        h = C.c_void_p(state._handle)
        assert c.lib.nes_pdas_free(C.byref(h), c.ptr) != 0
        state._handle = None
    if state._A is not None:
        state._A.free()
        state._A = None


def make_pdas(sf: StandardForm, scale=True, generated_seed=None) -> PdasState:
    """make-pdas (:75-133): clamp bounds, widen near-fixed variables, initial x from the unclamped
    bounds, z/w from the sign of c, y = 0, rows scaled by scale-constraints.

    `generated_seed`: the dense matrix is lpgen.dense_matrix(m, n, seed) and is regenerated on the
    device instead of uploaded (sf.A_dense may then be None); row scaling happens on the device."""
    nvars, ncons = sf.nvars, sf.ncons
    cvec = sf.c_dense()
    l = clamp(sf.l)
    u = clamp(sf.u)
    x = np.zeros(nvars)
    near = (u - l) < 1e-6
    l[near] -= 5e-7
    u[near] += 5e7  # sic (:93)
    sl, su = np.asarray(sf.l, dtype=np.float64), np.asarray(sf.u, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        delta = su - sl
        x_free = np.zeros(nvars)
        x_up = su - np.minimum(delta / 2, 1 + np.abs(su) * 0.1)
        x_lo = sl + np.minimum(delta / 2, 1 + np.abs(sl) * 0.1)
        x_mid = (sl + su) / 2
    x = np.where((sl < -1e10) & (su > 1e10), x_free,
                 np.where(sl < -1e6, x_up, np.where(su > 1e6, x_lo, x_mid)))
    z = np.where(cvec > 0, 1.0 + cvec, 1.0)
    w = np.where(cvec < 0, 1.0 - cvec, 1.0)
    b = np.array(sf.b, dtype=np.float64, copy=True)
    triplets, A_dense, A_generated = sf.A, sf.A_dense, None
    state_scale = None
    if generated_seed is not None:
        A_generated, A_dense, triplets = (generated_seed,), None, None
        state_scale = "device" if scale else None
    elif scale:
        if A_dense is not None:
            rowmax = np.abs(A_dense).max(axis=1)
            s = np.where(rowmax < 1e-6, 1.0, 1.0 / np.where(rowmax < 1e-6, 1.0, rowmax))
            A_dense = A_dense * s[:, None]
            b = b * s
        else:
            triplets, b = scale_constraints(triplets, b)
    return PdasState(nvars=nvars, ncons=ncons, c=cvec, triplets=triplets, b=b, l=l, u=u,
                     x0=x, y0=np.zeros(ncons), w0=w, z0=z, A_dense=A_dense, A_generated=A_generated,
                     _device_row_scale=state_scale)


def _ensure_device_row_scale(state: PdasState):
    if state._device_row_scale == "device":
        rs = state.A().scale_rows_maxabs()
        state.b = state.b * rs
        state._device_row_scale = None


def violation(state: PdasState):
    """violation (:135-150) + the scalars of one-pdas-iteration (:325-332).  The vectors
    (l, u, w.u, z.l, Ax-b, dual) stay on the device (state.get('l'|'u'|'p'|'d')); returns
    dict(pobj, dobj, violations[4], min_l, min_u)."""
    c = cholmod_common()
    out = (C.c_double * 8)()
    c.check(c.lib.nes_pdas_violation(state.handle(), out, c.ptr), "nes_pdas_violation")
    return {"pobj": out[0], "dobj": out[1], "violations": [out[2], out[3], out[4], out[5]],
            "min_l": out[6], "min_u": out[7]}


def direction(state: PdasState):
    """direction (:152-164) + pdas-step (:194-198): the Newton direction stays on the device
    (state.get('W'|'X'|'Y'|'Z')); returns alpha_max = min(box-step, pos-step w, pos-step z)."""
    c = cholmod_common()
    step = C.c_double(0.0)
    rc = c.check(c.lib.nes_pdas_newton_direction(state.handle(), C.byref(step), c.ptr),
                 "nes_pdas_newton_direction")
    if rc == nes.NES_DIV_BY_ZERO:   # the reference traps in scale-Z (SBCL DIVISION-BY-ZERO) after filter-Z
        raise ZeroDivisionError(c.error())
    if rc != 0:
        raise nes.NesError(f"solve-delta-y: Cholesky failed (status {rc}, minor {c.minor})")
    return step.value


def apply_step(state: PdasState, step):
    """apply-step (:200-207)."""
    c = cholmod_common()
    c.check(c.lib.nes_pdas_apply_step(state.handle(), float(step), c.ptr), "nes_pdas_apply_step")
    return state


def one_repair_iteration(state: PdasState):
    """one-repair-iteration (:268-288).  Returns (|g|, step)."""
    c = cholmod_common()
    out = (C.c_double * 2)()
    rc = c.check(c.lib.nes_pdas_repair(state.handle(), out, c.ptr), "nes_pdas_repair")
    if rc != 0:
        raise nes.NesError("cholesky-ls!: Cholesky failed")
    return out[0], out[1]


def recentre(state: PdasState):
    """The `repair` branch of one-pdas-iteration (:348-366)."""
    c = cholmod_common()
    out = (C.c_double * 2)()
    rc = c.check(c.lib.nes_pdas_recentre(state.handle(), out, c.ptr), "nes_pdas_recentre")
    if rc != 0:
        raise nes.NesError("primal-project: Cholesky failed")
    return out[0], out[1]


def one_pdas_iteration(state: PdasState, repair):
    """one-pdas-iteration (:319-383).  Returns (gap, dobj, step-or-None)."""
    v = violation(state)
    assert v["min_l"] > 0 and v["min_u"] > 0            # :323-324
    pobj, dobj, viol = v["pobj"], v["dobj"], v["violations"]
    gap = abs(pobj - dobj) / max(abs(pobj), abs(dobj), 1.0)   # :345-346
    primal_feasible = viol[0] < 1e-2                    # :333
    entry = {"pobj": pobj, "dobj": dobj, "violations": viol, "gap": gap}
    state.log.append(entry)
    if not primal_feasible:
        entry["branch"] = "repair"
        one_repair_iteration(state)
        return gap, dobj, None
    if repair:
        entry["branch"] = "recentre"
        recentre(state)
        return gap, dobj, None
    step = direction(state)
    entry["branch"], entry["step"] = "newton", step
    alpha = min(1.0, 0.9 * step)                        # :377-378
    assert 0.0 < alpha <= 1.0
    apply_step(state, alpha)
    return gap, dobj, step


def pdas(state: PdasState, max_iter=None, native_loop=False):
    """pdas (:385-396): iterate until the relative gap drops below 1e-4.
    Returns (dobj, gap, iterations).  Caller provides the with_cholmod() extent."""
    c = cholmod_common()
    try:
        state.handle()
        if native_loop:
            iters, obj, gap = C.c_int(0), C.c_double(0), C.c_double(0)
            rc = c.check(c.lib.nes_pdas_solve(state.handle(), max_iter or 0, C.byref(iters),
                                              C.byref(obj), C.byref(gap), c.ptr), "nes_pdas_solve")
            # NES_MAXITER: the loop ran out of iterations (the Lisp returns NIL there); the last iterate is
            # returned with state.converged = False so a caller cannot mistake it for a solution
            state.converged = rc == 0
            if rc == nes.NES_DIV_BY_ZERO:
                raise nes.NesError(f"pdas: {c.error()} (iteration {iters.value})")
            if rc not in (0, nes.NES_MAXITER):
                raise nes.NesError(f"pdas: Cholesky failed at iteration {iters.value}")
            state.final = {k: state.get(k) for k in "xywz"}
            return obj.value, gap.value, iters.value
        repair = False
        i = 0
        state.converged = False
        while max_iter is None or i < max_iter:
            i += 1
            viol, obj, step = one_pdas_iteration(state, repair)
            repair = step is not None and step < 1e-6       # :393
            if viol < 1e-4:                                  # :394
                state.converged = True
                break
        state.final = {k: state.get(k) for k in "xywz"}
        return obj, viol, i
    finally:
        free_pdas_A(state)
