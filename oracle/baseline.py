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


def affinity_cores():
    """CPUs this process may run on (not os.cpu_count(): a container / taskset may expose fewer)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """Give BLAS every core of the affinity mask, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its children).  Returns the thread count BLAS reports afterwards."""
    n = affinity_cores()
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass
    return host_threads()


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


# ---- bounded sample of a step too large to repeat on the host (BASELINE config 3) -------------------
# One step at m=32768, n=65536 is 8.2e13 flops: more than a minute on a 16-core host.  The timed sample
# is an exact 1/f cut of EVERY stage of that step at the full row count m (so BLAS sees the real operand
# shapes), not a smaller LP:
#   formation   rank-(n/f) update  M += (A_k diag s_k)(A_k diag s_k)'  of the full m x m normal matrix
#               (the k-th column slice of A; the step's dsyrk is the sum of f of these)       m^2 n / f
#   Cholesky    one right-looking step of width nb = m/f on a trailing matrix of order t, t chosen so
#               that nb^3/3 + t nb^2 + t^2 nb = m^3/(3f)  (the mean of the f panel steps): dpotrf of the
#               diagonal block, dtrsm of the panel, dsyrk of the trailing matrix
#   solves      forward + backward sweep over that panel of L                                2 m^2 / f
#   products    3 forward and 2 transposed dgemv over the m x (n/f) slice                    10 m n / f
def sample_sizes(m, n, f):
    ns, nb = n // f, m // f
    target = float(m) ** 3 / (3.0 * f)
    t = int((-nb * nb + (nb ** 4 - 4.0 * nb * (nb ** 3 / 3.0 - target)) ** 0.5) / (2.0 * nb))
    return ns, nb, min(t, m - nb)


def sample_flops(m, n, f):
    ns, nb, t = sample_sizes(m, n, f)
    return (float(m) * m * ns + nb ** 3 / 3.0 + float(t) * nb * nb + float(t) * t * nb
            + 2.0 * (nb * nb + 2.0 * t * nb) + 10.0 * m * ns)


def make_sample(m, n, f, seed=0):
    ns, nb, t = sample_sizes(m, n, f)
    rng = np.random.default_rng(seed)
    As = np.asfortranarray(rng.random((m, ns)))
    s = np.sqrt(0.1 + 10.0 * rng.random(ns))
    M = np.zeros((m, m), order="F")
    # trailing matrix of the Cholesky step: diagonally dominant so the diagonal block is positive definite
    T11 = np.asfortranarray(rng.random((nb, nb)))
    T11 = np.asfortranarray(np.tril(T11) + np.tril(T11, -1).T + nb * np.eye(nb))
    P0 = np.asfortranarray(rng.random((t, nb)))
    T22 = np.zeros((t, t), order="F")
    return {"m": m, "ns": ns, "nb": nb, "t": t, "As": As, "s": s, "M": M, "T11": T11, "P0": P0, "T22": T22,
            "g": rng.random(m), "v": rng.random(ns), "y": rng.random(nb + t)}


def sample_step(S):
    """One timed sample; returns seconds (operand restores between the stages are not timed)."""
    As, s, M = S["As"], S["s"], S["M"]
    t0 = time.perf_counter()
    B = As * s[None, :]
    M = blas.dsyrk(1.0, B, beta=1.0, c=M, lower=1, overwrite_c=1)
    S["M"] = M
    ax = As @ S["v"]
    g2 = S["g"] + As @ S["v"]
    ax2 = As @ S["v"]
    aty = As.T @ S["g"]
    r = As.T @ g2
    dt = time.perf_counter() - t0
    L11 = S["T11"].copy(order="F")
    P = S["P0"].copy(order="F")
    t0 = time.perf_counter()
    L11, info = scipy.linalg.lapack.dpotrf(L11, lower=1, overwrite_a=1)
    assert info == 0
    P = blas.dtrsm(1.0, L11, P, side=1, lower=1, trans_a=1, overwrite_b=1)
    S["T22"] = blas.dsyrk(-1.0, P, beta=1.0, c=S["T22"], lower=1, overwrite_c=1)
    y = S["y"].copy()
    nb = S["nb"]
    y[:nb] = blas.dtrsv(L11, y[:nb], lower=1)            # forward sweep over the panel
    y[nb:] -= P @ y[:nb]
    y[:nb] -= P.T @ y[nb:]                                # backward sweep
    y[:nb] = blas.dtrsv(L11, y[:nb], lower=1, trans=1)
    dt += time.perf_counter() - t0
    return dt


def time_sample(m, n, f, steps, warmup, seed=0):
    """Returns (seconds per sample list, flops per sample)."""
    S = make_sample(m, n, f, seed)
    times = []
    for i in range(warmup + steps):
        dt = sample_step(S)
        if i >= warmup:
            times.append(dt)
    return times, sample_flops(m, n, f)


def sample_description(m, n, f):
    ns, nb, t = sample_sizes(m, n, f)
    return (f"1/{f} of every stage of the m={m} n={n} step at full row count: dsyrk rank-{ns} update of the "
            f"{m}x{m} normal matrix + one right-looking Cholesky step (dpotrf {nb}, dtrsm {t}x{nb}, dsyrk "
            f"{t}x{t}x{nb}) + 2 sweeps over that panel + 5 dgemv over the {m}x{ns} slice")


# ---- config 4 / config 5 reference points (same "restated CPU path" status as everything above) ------
def sparse_step_cpu(rows, cols, vals, m, n, theta, b):
    """One sparse normal-equation step on the host: M = A diag(theta) A' (scipy.sparse), factor + solve with
    SuperLU in symmetric mode (the reference's CHOLMOD is not in this image).  Returns seconds per stage."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    A = sp.csc_matrix((vals, (rows, cols)), shape=(m, n))
    t0 = time.perf_counter()
    M = (A @ sp.diags(theta) @ A.T).tocsc()
    t1 = time.perf_counter()
    lu = spla.splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    t2 = time.perf_counter()
    x = lu.solve(b)
    t3 = time.perf_counter()
    res = float(np.linalg.norm(M @ x - b) / np.linalg.norm(b))
    return {"form_s": t1 - t0, "factor_s": t2 - t1, "solve_s": t3 - t2, "residual": res,
            "nnz_factor": int(lu.L.nnz)}


def batch_step_cpu(A, s, rhs):
    """Config 5 on the host: for every problem scale + dsyrk + dpotrf + dpotrs (LAPACK per problem, the way
    the reference would loop over its LPs).  A (B, m, n), s (B, n), rhs (B, m).  Returns seconds."""
    t0 = time.perf_counter()
    for k in range(A.shape[0]):
        Bk = np.asfortranarray(A[k] * s[k][None, :])
        M = blas.dsyrk(1.0, Bk, lower=1)
        cf = scipy.linalg.cho_factor(M, lower=True, overwrite_a=True, check_finite=False)
        scipy.linalg.cho_solve(cf, rhs[k], check_finite=False)
    return time.perf_counter() - t0
