"""Whole PDAS solve of BASELINE config 2 on the HOST with the oracle (restated reference CPU path), all cores:
the wall time bench.py's lp_solve.cpu_seconds estimate is checked against.  Run on the GPU box's host."""
import os, sys, time
ncpu = len(os.sched_getaffinity(0))
for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ[k] = str(ncpu)
import numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen
from oracle import pdas as opdas
from oracle import newton_solve as ons
import scipy.linalg.blas as blas
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = 2 * m
if os.environ.get("ORACLE_DSYRK", "1") == "1":
    # the cheapest CPU formulation of the same matrix (dsyrk instead of dgemm), as oracle/baseline.py times it
    def normal_matrix(A, s):
        B = np.asfortranarray(A * s[None, :])
        M = blas.dsyrk(1.0, B, lower=1)
        return np.tril(M) + np.tril(M, -1).T
    ons.normal_matrix = normal_matrix
t0 = time.perf_counter()
sf = lpgen.dense_lp(m, n, 0)
t1 = time.perf_counter()
st = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), sf.A_dense, sf.b, sf.l, sf.u)
t2 = time.perf_counter()
obj, gap, it = opdas.pdas(st, 300)
t3 = time.perf_counter()
print(f"host cores {ncpu}: m={m} n={n} generate {t1 - t0:.1f}s make-pdas {t2 - t1:.1f}s pdas {t3 - t2:.1f}s: {it} iterations, "
      f"dobj {obj:.9g}, gap {gap:.3g} -> {(t3 - t2) / it:.3f} s per iteration", flush=True)
