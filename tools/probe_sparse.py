"""Sparse path at BASELINE config-4 scale: symbolic analysis (host), numeric factorization, solve."""
import sys, time, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
bw = int(sys.argv[3]) if len(sys.argv) > 3 else 400
leaf = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t0 = time.perf_counter(); sf = lpgen.sparse_lp(m, n, nnz_per_col=10, bandwidth=bw, seed=0); print("gen %.2fs nnz %d" % (time.perf_counter() - t0, len(sf.A)))
with with_cholmod(device=0, timing=True) as c:
    c.lib.nes_set_ordering_leaf(c.ptr, leaf)
    t0 = time.perf_counter(); A = nes.Matrix.from_triplets(c, sf.A.row, sf.A.col, sf.A.value, m, n); print("upload %.2fs" % (time.perf_counter() - t0))
    A.scale(np.sqrt(0.1 + 10 * np.random.default_rng(0).random(n)))
    t0 = time.perf_counter(); L = nes.Factor(c, A); print("analyze %.2fs anz %.3g aatfl %.3g lnz %.3g fl %.3g mem %.2f GB" % (time.perf_counter() - t0, c.anz, c.aatfl, c.lnz, c.fl, c.memory_inuse / 1e9))
    b = np.random.default_rng(1).random(m)
    for i in range(3):
        c.timing_reset(); l0 = c.launches
        t0 = time.perf_counter(); ok = L.factorize(A); t1 = time.perf_counter(); x = L.solve(b); t2 = time.perf_counter()
        print("rep", i, ok, "factorize %.1f ms solve %.1f ms launches %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, c.launches - l0), {k: round(v[0], 2) for k, v in c.timing().items()}, flush=True)
    r = A.sdmult(A.sdmult(x, transpose=True)) - b
    print("solve residual", np.linalg.norm(r) / np.linalg.norm(b))
    L.free(); A.free()
