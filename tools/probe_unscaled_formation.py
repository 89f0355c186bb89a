import sys, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]); n = 2 * m
with with_cholmod(device=0, timing=True) as c:
    A = nes.Matrix.generate_dense(c, m, n, 0)
    L = nes.Factor(c, A)
    for scaled in (0, 1, 0, 1):
        if scaled: A.scale(np.sqrt(0.1 + 10 * np.random.default_rng(0).random(n)))
        else: A.unscale()
        c.timing_reset(); assert L.factorize(A); t = c.timing()
        print(f"m={m} scaled={scaled}: form {t['form'][0]:.2f} ms ({m*m*n/t['form'][0]/1e9:.2f} TFLOP/s) factor {t['factor'][0]:.1f}", flush=True)
    L.free(); A.free()
