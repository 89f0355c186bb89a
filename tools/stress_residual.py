"""Repeated factorizations at one size with the whole-matrix residual after each (single GPU)."""
import sys, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]); n = int(sys.argv[2]); reps = int(sys.argv[3])
with with_cholmod(device=0, timing=True) as c:
    A = nes.Matrix.generate_dense(c, m, n, 0)
    rng = np.random.default_rng(0)
    L = nes.Factor(c, A)
    out = []
    for r in range(reps):
        A.scale(np.sqrt(0.1 + 10 * rng.random(n)))
        c.timing_reset()
        assert L.factorize(A)
        out.append((round(c.timing()["factor"][0], 1), L.residual(A)))
    print(f"m={m} n={n}:", " ".join(f"{t}ms/{r:.1e}" for t, r in out), flush=True)
    L.free(); A.free()
