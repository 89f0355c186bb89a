"""CPU reference point for BASELINE config 4 (not the reference's CHOLMOD, which is not in this image):
form M = A diag(theta) A' with scipy.sparse and factor/solve it with SuperLU (scipy.sparse.linalg.splu,
symmetric mode, COLAMD/MMD ordering) -- one normal-equation step on the host cores."""
import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
bw = int(sys.argv[3]) if len(sys.argv) > 3 else 400
sf = lpgen.sparse_lp(m, n, nnz_per_col=10, bandwidth=bw, seed=0)
A = sp.csc_matrix((sf.A.value, (sf.A.row, sf.A.col)), shape=(m, n))
theta = 0.1 + 10 * np.random.default_rng(0).random(n)
b = np.random.default_rng(1).random(m)
t0 = time.perf_counter(); M = (A @ sp.diags(theta) @ A.T).tocsc(); t1 = time.perf_counter()
lu = spla.splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True)); t2 = time.perf_counter()
x = lu.solve(b); t3 = time.perf_counter()
print(f"m={m} n={n}: form {t1-t0:.2f}s factor {t2-t1:.2f}s solve {t3-t2:.3f}s nnz(L+U) {lu.L.nnz + lu.U.nnz:.3g} "
      f"residual {np.linalg.norm(M @ x - b) / np.linalg.norm(b):.2e}")
