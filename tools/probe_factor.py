"""Short driver for profiling: one formation + Cholesky + solve at m x n (default BASELINE config 2)."""
import sys, time, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * m
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
with with_cholmod(device=0, timing=True) as c:
    A = nes.Matrix.generate_dense(c, m, n, 0)
    A.scale(np.sqrt(0.1 + 10 * np.random.default_rng(0).random(n)))
    L = nes.Factor(c, A)
    b = np.random.default_rng(1).random(m)
    for i in range(reps):
        c.timing_reset()
        t0 = time.perf_counter(); ok = L.factorize(A); x = L.solve(b); t1 = time.perf_counter()
        print("rep", i, ok, "%.2f ms" % ((t1 - t0) * 1e3), {k: round(v[0], 3) for k, v in c.timing().items()}, flush=True)
    r = A.sdmult(A.sdmult(x, transpose=True)) - b   # (A s)(A s)' x - b through the GEMV kernels
    print("solve residual", np.linalg.norm(r) / np.linalg.norm(b))
    L.free(); A.free()
