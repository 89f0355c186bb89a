"""Oracle (test infrastructure): the timed CPU baseline -- "restated reference CPU path".

One IPM normal-equation step on the host, with the same algebra as oracle/newton_solve.py but
using the cheapest CPU formulation (so GPU/CPU ratios are not inflated by the reference's
2mn^2 diagonal dgemm, newton-solve.lisp:112-116, or CHOLMOD's scalar A*A' assembly):
    B = A diag(s)            O(mn)          (cholmod_scale, sparse-cholesky.lisp:461-473)
    M = B B'                 dsyrk, m^2 n   (cholmod_factorize forms B B', :419)
    L L' = M                 dpotrf, m^3/3
    dy = L'^-1 L^-1 g        dpotrs, 2 m^2
    A v, A' dy + the two products of `violation`   5 gemv, 10 mn
All BLAS/LAPACK calls go to SciPy's bundled OpenBLAS with every host thread it wants.
Only bench.py's cpu_baseline / --impl reference legs call this.
"""
from __future__ import annotations

import os
import time

import numpy as np
import scipy.linalg
import scipy.linalg.blas as blas


def flops(m, n):
    """Algorithmic flops of one dense IPM iteration."""
    return float(m) * m * n + float(m) ** 3 / 3.0 + 2.0 * m * m + 10.0 * m * n


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max((p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"),
                default=None)
        if n:
            return int(n)
    except Exception:
        pass
    return os.cpu_count() or 1


def make_inputs(m, n, seed=0):
    rng = np.random.default_rng(seed)
    A = np.asfortranarray(rng.random((m, n)))
    A[np.arange(min(m, n)), np.arange(min(m, n))] += 1.0
    s = np.sqrt(0.1 + 10.0 * rng.random(n))
    g = rng.random(m)
    v = rng.random(n)
    return A, s, g, v


def normal_eq_step(A, s, g, v):
    """Returns dy; raises LinAlgError when M is not positive definite."""
    x = v.copy()
    ax = A @ x                      # violation: A x
    aty = A.T @ g                   # violation: A' y
    g2 = g + A @ v                  # g + A(l'h3 - f')   (two fused forward products)
    B = A * s[None, :]
    M = blas.dsyrk(1.0, B, lower=1)
    c = scipy.linalg.cho_factor(M, lower=True, overwrite_a=True, check_finite=False)
    dy = scipy.linalg.cho_solve(c, g2, check_finite=False)
    r = A.T @ dy                    # solve-delta-z
    return dy, ax, aty, r


def time_steps(m, n, steps, warmup, seed=0):
    """Returns (seconds per step list, flops per step)."""
    A, s, g, v = make_inputs(m, n, seed)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        normal_eq_step(A, s, g, v)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, flops(m, n)


def pick_sample(m, n, budget_s=25.0):
    """Largest (m', n') with the same aspect ratio whose step is estimated to fit the budget."""
    t, f = time_steps(1024, 1024 * n // m, 1, 1)
    rate = f / t[0]
    mm, nn = m, n
    while flops(mm, nn) / rate > budget_s and mm > 1024:
        mm //= 2
        nn //= 2
    return mm, nn, rate
