"""Small sparse factor + solve (debug driver for compute-sanitizer runs)."""
import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
from oracle import newton_solve as ons
m = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rng = np.random.default_rng(m)
A = sp.csc_matrix(ons.random_sparse_matrix(rng, m, n, 0.05 if m <= 60 else 0.01)); A.sort_indices()
s = np.sqrt(0.1 + 10 * rng.random(n)); b = rng.random(m)
with with_cholmod(device=0) as c:
    Ad = nes.Matrix.from_csc(c, A.indptr, A.indices, A.data, m, n); Ad.scale(s)
    L = nes.Factor(c, Ad)
    print("factorize", L.factorize(Ad), flush=True)
    x = L.solve(b)
    M = ons.normal_matrix(A, s)
    print("residual", np.linalg.norm(M @ x - b) / np.linalg.norm(b))
    L.free(); Ad.free()
