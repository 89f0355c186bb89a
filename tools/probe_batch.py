"""BASELINE config 5: 1024 dense LPs of m=256 (n=512) -- batched normal-equation solves and affine scaling."""
import sys, time, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import batched
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m, n = 256, 512
rng = np.random.default_rng(0)
A = rng.random((B, m, n)); A[:, np.arange(m), np.arange(m)] += 1.0
xs = 0.1 + 10 * rng.random((B, n)); b = np.einsum("bmn,bn->bm", A, xs)
ys = rng.uniform(-1, 1, (B, m)); zs = 0.1 + 10 * rng.random((B, n)); c = np.einsum("bmn,bm->bn", A, ys) + zs
l = np.zeros((B, n)); u = np.full((B, n), np.inf); x0 = np.ones((B, n))
F = (m * m * n + m ** 3 / 3 + 2 * m * m + 10 * m * n) * B
with with_cholmod(device=0, timing=True) as cm:
    bt = batched.Batch(A, c=c, b=b, l=l, u=u, x=x0)
    s = np.sqrt(0.1 + 10 * rng.random((B, n))); rhs = rng.random((B, m))
    for i in range(3):
        cm.timing_reset(); t0 = time.perf_counter(); x, st = bt.normal_solve(s, rhs); dt = time.perf_counter() - t0
        print("normal_solve %.2f ms (host wall incl. PCIe)" % (dt * 1e3), {k: round(v[0], 3) for k, v in cm.timing().items()}, "fail", int(st.sum()), flush=True)
    b0 = 7
    M = (A[b0] * s[b0] ** 2) @ A[b0].T
    print("residual problem 7:", np.linalg.norm(M @ x[b0] - rhs[b0]) / np.linalg.norm(rhs[b0]))
    cm.timing_reset(); l0 = cm.launches; t0 = time.perf_counter()
    obj, xx, res, iters = bt.affine_scaling(3000)
    dt = time.perf_counter() - t0
    tm = cm.timing(); nit = int(iters.max())
    dev = sum(v[0] for v in tm.values())
    print("affine: %.2f s wall, max iters %d, mean iters %.1f, launches %d" % (dt, nit, iters.mean(), cm.launches - l0))
    print("stage ms", {k: round(v[0], 1) for k, v in tm.items()}, "device total %.1f ms -> %.1f GFLOP/s per batched iteration (stage time)" % (dev, F * nit / (dev * 1e-3) / 1e9))
    print("objective spread", obj.min(), obj.max(), "max residual", res.max())
    bt.free()
