import sys, time, ctypes as C, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m, n = 8192, 16384
with with_cholmod(device=0) as c:
    A = nes.Matrix.generate_dense(c, m, n, 0)
    L = nes.Factor(c, A)
    rng = np.random.default_rng(0)
    v = [0.1 + 10 * rng.random(n) for _ in range(4)] + [rng.random(n), rng.random(n), rng.random(m), rng.random(n)]
    ins = [nes.vec(x) for x in v]
    outs = [np.empty(n), np.empty(n), np.empty(m), np.empty(n)]
    for i in range(8):
        t0 = time.perf_counter()
        rc = c.lib.nes_kkt_newton(A.ptr, L.ptr, 0, *[p for _, p in ins], *[o.ctypes.data_as(nes._dp) for o in outs], c.ptr)
        t1 = time.perf_counter()
        print("kkt call", i, rc, (t1 - t0) * 1e3, "ms", flush=True)
    for i in range(4):
        t0 = time.perf_counter(); B = A.copy(); t1 = time.perf_counter(); B.scale(np.ones(n)); t2 = time.perf_counter(); B.free(); t3 = time.perf_counter()
        print("copy %.3f scale %.3f free %.3f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), flush=True)
    L.free(); A.free()
