"""solve-kkt-newton (newton-solve.lisp:139-154 dense, sparse-newton-solve.lisp:150-168 sparse).

The reference composes ~12 matlisp calls around one CHOLMOD solve; here the whole reduction runs
on the device behind one C-ABI call (nes_kkt_newton): fused elementwise pre/post passes, the fused
scale+SYRK formation, the DMMA Cholesky and the triangular solves.  Inputs are not destroyed.
"""
from __future__ import annotations

import numpy as np

from . import nes
from .sparse_cholesky import cholmod_common


def solve_kkt_newton(l, u, w, z, A, e, f, g, h, filters=None, factor=None):
    """Returns (dw, dx, dy, dz).  `A` is a nes.Matrix (dense or sparse).  `filters` defaults to the
    reference's behaviour per file: off for dense A (newton-solve.lisp), on for sparse A
    (sparse-newton-solve.lisp filter-U / filter-Z).  Raises if the Cholesky fails, like the Lisp's
    (the real-matrix ...) / (assert dy) on the NIL (sparse-newton-solve.lisp:125, 139)."""
    c = cholmod_common()
    if filters is None:
        filters = not A.is_dense
    m, n = A.shape
    ins = [nes.vec(v) for v in (l, u, w, z, e, f, g, h)]
    for arr, want in zip(ins, (n, n, n, n, n, n, m, n)):
        assert len(arr[0]) == want
    dw, dx, dz, dy = np.empty(n), np.empty(n), np.empty(n), np.empty(m)
    rc = c.lib.nes_kkt_newton(A.ptr, factor.ptr if factor is not None else None, 1 if filters else 0,
                              *[p for _, p in ins],
                              dw.ctypes.data_as(nes._dp), dx.ctypes.data_as(nes._dp),
                              dy.ctypes.data_as(nes._dp), dz.ctypes.data_as(nes._dp), c.ptr)
    c.check(rc, "nes_kkt_newton")
    if rc == nes.NES_DIV_BY_ZERO:   # the reference traps in scale-Z (SBCL DIVISION-BY-ZERO) after filter-Z
        raise ZeroDivisionError(c.error())
    if rc != 0:
        raise nes.NesError(f"solve-delta-y: Cholesky failed (status {rc}, minor {c.minor})")
    return dw, dx, dy, dz


def test_kkt_solve(l, u, w, z, A, e, f, g, h, A_host, ord=2, filters=None):
    """test-kkt-solve (newton-solve.lisp:166-182; sparse-newton-solve.lisp:180-198 with ord=inf):
    residuals of the four un-reduced block rows.  `A_host` is A as a NumPy / SciPy matrix."""
    dw, dx, dy, dz = solve_kkt_newton(l, u, w, z, A, e, f, g, h, filters=filters)
    nrm = lambda v: np.linalg.norm(v, ord)
    return (nrm(u * dw - w * dx - e), nrm(z * dx + l * dz - f),
            nrm(np.asarray(A_host @ dx).ravel() - g),
            nrm(np.asarray(A_host.T @ dy).ravel() + dz - dw - h))
