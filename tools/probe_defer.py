"""Sweep of the deferred-formation knobs of dense_chol.cu (NES_CHOL_DEFER columns, NES_CHOL_RESERVE SMs):
form + factor + solve time of one normal-equation step at m x n (default BASELINE config 2)."""
import os, sys, time, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * m
configs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]] or [(0, 8), (4096, 8)]
with with_cholmod(device=0, timing=True) as c:
    A = nes.Matrix.generate_dense(c, m, n, 0)
    A.scale(np.sqrt(0.1 + 10 * np.random.default_rng(0).random(n)))
    b = np.random.default_rng(1).random(m)
    for defer, reserve in configs:
        os.environ["NES_CHOL_DEFER"] = str(defer)
        os.environ["NES_CHOL_RESERVE"] = str(reserve)
        L = nes.Factor(c, A)
        best = None
        for i in range(4):
            c.timing_reset()
            c.mark_begin()
            ok = L.factorize(A)
            ms = c.mark_end()
            t = {k: round(v[0], 3) for k, v in c.timing().items()}
            if i and (best is None or ms < best[0]):
                best = (ms, t)
        x = L.solve(b)
        r = A.sdmult(A.sdmult(x, transpose=True)) - b
        print(f"defer {defer:5d} reserve {reserve:2d}: form+factor {best[0]:.3f} ms  form {best[1]['form']:.3f} "
              f"factor {best[1]['factor']:.3f}  upfront flops {c.form_flops:.4e}  ok {ok} "
              f"residual {np.linalg.norm(r) / np.linalg.norm(b):.2e}", flush=True)
        L.free()
    A.free()
