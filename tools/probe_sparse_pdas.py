"""BASELINE config 4 end to end: synthetic sparse LP (m=100k, n=250k, ~10 nnz/col) -> make-pdas -> pdas with
the sparse Newton solve on the GPU.  Prints iterations and wall seconds (analysis included)."""
import sys, time, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, pdas
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
bw = int(sys.argv[3]) if len(sys.argv) > 3 else 400
maxit = int(sys.argv[4]) if len(sys.argv) > 4 else 300
t0 = time.perf_counter(); sf = lpgen.sparse_lp(m, n, nnz_per_col=10, bandwidth=bw, seed=0); tg = time.perf_counter() - t0
with with_cholmod(device=0, timing=True) as c:
    t0 = time.perf_counter(); st = pdas.make_pdas(sf); tm = time.perf_counter() - t0
    l0 = c.launches
    t0 = time.perf_counter(); obj, gap, it = pdas.pdas(st, maxit, native_loop=True); ts = time.perf_counter() - t0
    print(f"generate {tg:.2f}s make-pdas {tm:.2f}s pdas {ts:.2f}s: {it} iterations, objective {obj:.9g}, gap {gap:.3g}, "
          f"{c.launches - l0} launches; device stage ms {({k: round(v[0], 1) for k, v in c.timing().items()})}", flush=True)
