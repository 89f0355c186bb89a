"""One formation + factorization at BASELINE config 3 (or argv sizes): the command ncu profiles
(ncu -k regex:dmma_nt_kernel -c 1 captures the fused scale+SYRK launch)."""
import sys, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * m
with with_cholmod(device=0, timing=True) as c:
    A = nes.Matrix.generate_dense(c, m, n, 0)
    A.scale(np.sqrt(0.1 + 10 * np.random.default_rng(0).random(n)))
    L = nes.Factor(c, A)
    assert L.factorize(A)
    t = c.timing()
    res = L.residual(A) if len(sys.argv) > 3 else float("nan")
    print(f"m={m} n={n}: residual {res:.2e} form {t['form'][0]:.2f} ms ({m * m * n / t['form'][0] / 1e9:.2f} TFLOP/s)  factor {t['factor'][0]:.2f} ms", flush=True)
    L.free(); A.free()
