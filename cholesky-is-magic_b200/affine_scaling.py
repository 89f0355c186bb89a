"""Primal affine scaling driver (affine-scaling.lisp), host-side mirror over the C ABI.

    min c'x  s.t.  Ax = b,  l <= x <= u

Function names, constants and control flow follow the Lisp; x and every temporary live on the GPU in
a `nes_affine` handle, the symbolic analysis happens once (affine-scaling.lisp:270-271) and each
iteration refactorizes numerically (solve-sparse-recycle).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import nes
from .sparse_cholesky import cholmod_common, make_sparse_from_triplet_vector
from .standard_form import StandardForm, Triplets

MAX_SLACK = 1e8   # *max-slack* (:118)
GAMMA = 0.9       # *gamma* (:135)


@dataclass
class AffineScalingState:
    """affine-scaling-state (:1-10)."""
    nvars: int
    ncons: int
    x0: np.ndarray
    c: np.ndarray
    triplets: Triplets | None
    b: np.ndarray
    l: np.ndarray
    u: np.ndarray
    A_dense: np.ndarray | None = None
    _A: nes.Matrix | None = None
    _handle: int | None = None
    log: list = field(default_factory=list)

    def A(self):
        """affine-A (:12-27)."""
        if self._A is None:
            c = cholmod_common()
            if self.A_dense is not None:
                self._A = nes.Matrix.from_dense(c, self.A_dense)
            else:
                self._A = make_sparse_from_triplet_vector(self.ncons, self.nvars, self.triplets)
        return self._A

    def handle(self):
        if self._handle is None:
            c = cholmod_common()
            keep = [nes.vec(v)[0] for v in (self.c, self.b, self.l, self.u, self.x0)]
            h = c.lib.nes_affine_create(self.A().ptr, *[k.ctypes.data_as(nes._dp) for k in keep], c.ptr)
            if not h:
                raise nes.NesError(f"nes_affine_create failed: {c.error()}")
            self._handle = h
            # the counters the Lisp prints after cholmod_analyze (:273-279)
            self.counters = {"anz": c.anz, "aatfl": c.aatfl, "lnz": c.lnz, "fl": c.fl}
        return self._handle

    def get(self, which):
        c = cholmod_common()
        out = np.empty(self.ncons if which == "r" else self.nvars)
        c.check(c.lib.nes_affine_get(self.handle(), ord(which), out.ctypes.data_as(nes._dp), c.ptr), "nes_affine_get")
        return out

    x = property(lambda self: self.get("x"))


def free_affine_A(state: AffineScalingState):
    """free-affine-A (:44-50) + free-sparse-state."""
    c = cholmod_common()
    if state._handle is not None:
        h = C.c_void_p(state._handle)
        assert c.lib.nes_affine_free(C.byref(h), c.ptr) != 0
        state._handle = None
    if state._A is not None:
        state._A.free()
        state._A = None


def make_affine_state(sf: StandardForm) -> AffineScalingState:
    """make-affine-state (:52-90): bounds are NOT clamped; near-fixed variables are widened; initial x
    by the 1e10 rules (lower-bounded case: 1 + |l| * 1.0, :74-75)."""
    n = sf.nvars
    l = np.array(sf.l, dtype=np.float64, copy=True)
    u = np.array(sf.u, dtype=np.float64, copy=True)
    near = (u - l) < 1e-6
    l[near] -= 5e-7
    u[near] += 5e7
    with np.errstate(invalid="ignore"):
        delta = u - l
        x = np.where((l < -1e10) & (u > 1e10), 0.0,
                     np.where(l < -1e10, u - np.minimum(delta / 2, 1 + np.abs(u) * 0.1),
                              np.where(u > 1e10, l + np.minimum(delta / 2, 1 + np.abs(l) * 1.0), (l + u) / 2)))
    return AffineScalingState(nvars=n, ncons=sf.ncons, x0=x, c=sf.c_dense(), triplets=sf.A,
                              b=np.array(sf.b, dtype=np.float64, copy=True), l=l, u=u, A_dense=sf.A_dense)


def residual(state):
    """residual (:209-213): returns (|b - Ax|_2, c'x); the vector stays on the device ('r')."""
    c = cholmod_common()
    out = (C.c_double * 2)()
    c.check(c.lib.nes_affine_residual(state.handle(), out, c.ptr), "nes_affine_residual")
    return out[0], out[1]


def one_repair_iteration(state):
    """one-repair-iteration (:226-243) on the last residual.  Returns (state, t)."""
    c = cholmod_common()
    out = (C.c_double * 2)()
    rc = c.check(c.lib.nes_affine_repair(state.handle(), out, c.ptr), "nes_affine_repair")
    if rc != 0:
        raise nes.NesError("cholesky-ls!: Cholesky failed")
    return state, True


def one_affine_scaling_iteration(state, centering=False):
    """one-affine-scaling-iteration (:165-207).  Returns (state, continue)."""
    c = cholmod_common()
    out = (C.c_double * 5)()
    rc = c.check(c.lib.nes_affine_direction(state.handle(), 1 if centering else 0, out, c.ptr),
                 "nes_affine_direction")
    if rc != 0:                       # " singular " (:178-181)
        return state, False
    step, norm_g, norm_dg, descent, min_slack = out
    assert min_slack > 0              # slack assert (:144)
    if step > 1e10:
        raise ArithmeticError("Unbounded problem")      # (:187-188)
    if not centering:
        if norm_dg < min(1e-6, 1e-8 * state.nvars) or descent > 0:
            return state, False
        if step * norm_g < 1e-6 or descent > 0:
            return one_affine_scaling_iteration(state, centering=True)
    c.check(c.lib.nes_affine_apply(state.handle(), step, c.ptr), "nes_affine_apply")
    return state, True


def one_iteration(state, centering=False):
    """one-iteration (:245-263)."""
    norm, obj = residual(state)
    if norm > 1e-6 * state.ncons:
        state.log.append(("repair", norm))
        return one_repair_iteration(state)
    state.log.append(("recenter" if centering else "optimize", obj))
    return one_affine_scaling_iteration(state, centering=centering)


def affine_scaling(state, max_iter=100000, native_loop=False):
    """affine-scaling (:265-297).  Returns (c'x, x, |residual|, iterations)."""
    c = cholmod_common()
    try:
        state.handle()
        if native_loop:
            iters, obj, res = C.c_int(0), C.c_double(0), C.c_double(0)
            rc = c.check(c.lib.nes_affine_solve(state.handle(), max_iter, C.byref(iters), C.byref(obj),
                                                C.byref(res), c.ptr), "nes_affine_solve")
            state.converged = rc == 0
            if rc not in (0, nes.NES_MAXITER):
                raise nes.NesError("affine-scaling: Cholesky failed")
            return obj.value, state.x, res.value, iters.value
        i = 0
        while i < max_iter:
            _, cont = one_iteration(state, (i + 1) % 16 == 0)
            norm, obj = residual(state)
            if not (cont or norm > 1e-6 * state.ncons):
                state.converged = True
                return obj, state.x, norm, i + 1
            i += 1
        norm, obj = residual(state)
        state.converged = False
        return obj, state.x, norm, i
    finally:
        free_affine_A(state)
