"""Synthetic standard-form LPs for the BASELINE configurations.

Dense: A = U(0,1)^{m x n} + eye(m,n)  (generator of newton-solve.lisp:194-195), an interior point
x* = 0.1 + 10 U(0,1) (distribution of :184-185), b = A x*, y* = U(-1,1), z* = 0.1 + 10 U(0,1),
c = A'y* + z*  => primal and dual strictly feasible, bounded.  Bounds lo = 0, hi = +inf.

The entries come from a counter-based hash so the GPU can regenerate the same matrix without a PCIe
upload (csrc/nes_matrix.cu generate_dense_kernel computes exactly `dense_entry`).
"""
from __future__ import annotations

import numpy as np

from .standard_form import StandardForm, Triplets

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def dense_entry(seed, i, j):
    """U(0,1) value of A[i, j] before the identity is added (vectorised over i, j)."""
    with np.errstate(over="ignore"):
        i = np.asarray(i, dtype=np.uint64)
        j = np.asarray(j, dtype=np.uint64)
        s = _mix64(np.uint64(seed))
        h = _mix64(s ^ (i * np.uint64(0x100000001B3) + j * np.uint64(0xC2B2AE3D27D4EB4F)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def dense_matrix(m, n, seed):
    i = np.arange(m, dtype=np.uint64)[:, None]
    j = np.arange(n, dtype=np.uint64)[None, :]
    A = dense_entry(seed, i, j)
    k = np.arange(min(m, n))
    A[k, k] += 1.0
    return A


def aux_vectors(m, n, seed):
    """x*, y*, z* of the feasible construction (independent stream from A's hash)."""
    rng = np.random.default_rng([int(seed), 0x5EED])
    xs = 0.1 + 10.0 * rng.random(n)
    ys = rng.uniform(-1.0, 1.0, m)
    zs = 0.1 + 10.0 * rng.random(n)
    return xs, ys, zs


def dense_lp(m, n, seed=0, A=None):
    """Standard-form dense LP.  `A` may be passed in (e.g. computed elsewhere) to skip generation."""
    if A is None:
        A = dense_matrix(m, n, seed)
    xs, ys, zs = aux_vectors(m, n, seed)
    b = A @ xs
    cvec = A.T @ ys + zs
    return StandardForm(
        nvars=n, ncons=m, c=[(i, float(v)) for i, v in enumerate(cvec)], A=None, b=b,
        l=np.zeros(n), u=np.full(n, np.inf), type=[None] * m, initial_vars=n, A_dense=A)


def sparse_lp(m, n, nnz_per_col=10, bandwidth=None, seed=0):
    """Structured sparse LP for config 4: column j has a forced entry on row j mod m (the unit-pattern
    diagonal of sparse-newton-solve.lisp:234, which guarantees full row rank) plus nnz_per_col-1
    entries inside a band of `bandwidth` rows around it; values 1 + U(0,1) (:236).  A banded pattern
    keeps nnz(L) of A diag(theta) A' at O(m * bandwidth) instead of filling in completely."""
    rng = np.random.default_rng([int(seed), 0x5A5A])
    if bandwidth is None:
        bandwidth = max(4 * nnz_per_col, 32)
    bandwidth = min(bandwidth, m)
    cols = np.repeat(np.arange(n, dtype=np.int64), nnz_per_col)
    centre = (np.arange(n, dtype=np.int64) * m) // n if n >= m else np.arange(n, dtype=np.int64)
    centre = np.where(np.arange(n) < m, np.arange(n), centre)  # first m columns hit the diagonal
    off = rng.integers(-(bandwidth // 2), bandwidth // 2 + 1, size=(n, nnz_per_col))
    off[:, 0] = 0
    rows = np.clip(centre[:, None] + off, 0, m - 1).reshape(-1)
    vals = 1.0 + rng.random(n * nnz_per_col)
    # merge duplicates the way cholmod_triplet_to_sparse does (sum)
    key = cols * m + rows
    order = np.argsort(key, kind="stable")
    key, rows, cols, vals = key[order], rows[order], cols[order], vals[order]
    first = np.concatenate(([True], key[1:] != key[:-1]))
    seg = np.cumsum(first) - 1
    vals = np.bincount(seg, weights=vals)
    rows, cols = rows[first], cols[first]
    xs, ys, zs = aux_vectors(m, n, seed)
    b = np.bincount(rows, weights=vals * xs[cols], minlength=m)
    cvec = np.bincount(cols, weights=vals * ys[rows], minlength=n) + zs
    return StandardForm(
        nvars=n, ncons=m, c=[(i, float(v)) for i, v in enumerate(cvec)],
        A=Triplets(rows.astype(np.int32), cols.astype(np.int32), vals), b=np.asarray(b).ravel(),
        l=np.zeros(n), u=np.full(n, np.inf), type=[None] * m, initial_vars=n)
