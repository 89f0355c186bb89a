"""Host-side mirror of the marshalling + solver glue of sparse-cholesky.lisp:344-614.

Same names and behaviour as the Lisp (hyphens -> underscores, `!` dropped, `*` spelled `star`);
every function is a few C-ABI calls.  A dynamically bound ``*cholmod-common*`` is played by the
module-level `_common` stack managed by `with_cholmod`.
"""
from __future__ import annotations

import contextlib

import numpy as np

from . import nes

_common_stack: list = []


def cholmod_common() -> nes.Common:
    """*cholmod-common* (sparse-cholesky.lisp:344)."""
    if not _common_stack:
        raise nes.NesError("*cholmod-common* is unbound: call inside with_cholmod()")
    return _common_stack[-1]


@contextlib.contextmanager
def with_cholmod(device=None, timing=False):
    """with-cholmod (sparse-cholesky.lisp:400-406): make-default-common ... free-common."""
    common = nes.Common(device=device, timing=timing)
    _common_stack.append(common)
    try:
        yield common
    finally:
        _common_stack.pop()
        common.close()


def make_sparse_from_triplet_vector(nrow, ncol, triplets):
    """make-sparse-from-triplet-vector (sparse-cholesky.lisp:433-459)."""
    return nes.Matrix.from_triplets(cholmod_common(), triplets.row, triplets.col, triplets.value,
                                    nrow, ncol)


def solve_dense(A, b):
    """solve-dense (sparse-cholesky.lisp:409-431): x with (A A') x = b for a dense m x n A, or None
    when the factorization status is non-zero."""
    c = cholmod_common()
    A = np.asfortranarray(A, dtype=np.float64)
    m, n = A.shape
    b, bp = nes.vec(b)
    x = np.empty(m)
    rc = c.lib.nes_solve_dense(A.ctypes.data_as(nes._dp), m, n, bp, x.ctypes.data_as(nes._dp), c.ptr)
    c.check(rc, "nes_solve_dense")
    return None if rc != 0 else x


def scale_sparse_(sparse, scale):
    """scale-sparse! (sparse-cholesky.lisp:461-473): columns of `sparse` times `scale`, in place."""
    return sparse.scale(scale)


def scale_sparse(sparse, scale):
    """scale-sparse (sparse-cholesky.lisp:475-477): scaled copy."""
    return sparse.copy().scale(scale)


class SolveSparseState:
    """solve-sparse-state (sparse-cholesky.lisp:479-484): the recycled factor + workspaces."""

    def __init__(self):
        self.factor = None


def free_sparse_state(state):
    """free-sparse-state (sparse-cholesky.lisp:486-504)."""
    if state.factor is not None:
        state.factor.free()
        state.factor = None


def solve_sparse_one_shot(As, b):
    """solve-sparse-one-shot (sparse-cholesky.lisp:506-522): analyze, factorize, solve, free."""
    c = cholmod_common()
    factor = nes.Factor(c, As)
    try:
        if not factor.factorize(As):
            return None
        return factor.solve(b)
    finally:
        factor.free()


def solve_sparse_recycle(As, b, state, factorized):
    """solve-sparse-recycle (sparse-cholesky.lisp:524-560): symbolic factor and workspaces reused;
    `factorized` skips the numeric refactorization."""
    c = cholmod_common()
    if state.factor is None:
        state.factor = nes.Factor(c, As)
    if not factorized:
        if not state.factor.factorize(As):
            return None
    return state.factor.solve(b)


def solve_sparse(As, b, state=None, factorized=False):
    """solve-sparse (sparse-cholesky.lisp:562-565)."""
    if state is not None:
        return solve_sparse_recycle(As, b, state, factorized)
    return solve_sparse_one_shot(As, b)


def sparse_m_star(sparse, x, transpose=False, y=None, alpha=1.0, beta=None):
    """sparse-m* (sparse-cholesky.lisp:567-614): alpha op(A) x + beta y (fresh result)."""
    return sparse.sdmult(x, transpose=transpose, y=y, alpha=alpha, beta=beta)
