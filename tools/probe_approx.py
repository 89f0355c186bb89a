"""APPROX (approx.lisp) at BASELINE config-4 scale on the GPU: time per iteration and achieved HBM
bandwidth.  One iteration = two value-&-gradient evaluations = 4 passes over the stacked constraint matrix K."""
import sys, time, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, approx as gap
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 500
sf = lpgen.sparse_lp(m, n, nnz_per_col=10, bandwidth=400, seed=0)
import os
with with_cholmod(device=0, timing=bool(int(os.environ.get("NES_PROBE_TIMING", "0")))) as c:
    t0 = time.perf_counter(); st = gap.make_approx(sf); tm = time.perf_counter() - t0
    nnzK = st.K.nnz
    N, R = st.nvars, len(st.rhs)
    gap.approx(st, 20)
    c.timing_reset(); l0 = c.launches
    t0 = time.perf_counter(); z, it, rs, stats = gap.approx(st, iters); dt = time.perf_counter() - t0
    # bytes per iteration: 2 gradients x (K by rows + K by columns: 12 B per entry each) + vector passes
    per_grad = 2 * nnzK * 12 + (3 * R + 6 * N) * 8
    per_iter = 2 * per_grad + 14 * N * 8
    print(f"make-approx {tm:.2f}s: N={N} R={R} nnz(K)={nnzK}; {it} iterations in {dt*1e3:.1f} ms = {dt/it*1e6:.1f} us/iteration, "
          f"{(c.launches - l0)/it:.1f} launches/iteration, {per_iter/1e6:.1f} MB/iteration -> {per_iter*it/dt/1e9:.0f} GB/s; "
          f"restarts {rs}, value {stats[3]:.4g}, |pg| {stats[1]:.3g}; device ms {({k: round(v[0], 1) for k, v in c.timing().items()})}")
    st.free()
