"""PDAS on structured sparse LPs of growing size (GPU sparse path): iterations, objective, failures."""
import sys, time, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, pdas, nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
dbound = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
for (m, n, bw) in ((2000, 5000, 400), (8000, 20000, 400), (30000, 75000, 400), (100000, 250000, 400)):
    sf = lpgen.sparse_lp(m, n, nnz_per_col=10, bandwidth=bw, seed=0)
    with with_cholmod(device=0, timing=True) as c:
        c.set("dbound", dbound)
        st = pdas.make_pdas(sf)
        t0 = time.perf_counter()
        try:
            obj, gap, it = pdas.pdas(st, 400, native_loop=True)
            print(f"dbound {dbound:g} m={m}: {it} iterations obj {obj:.12g} gap {gap:.3g} in {time.perf_counter() - t0:.2f}s "
                  f"{({k: round(v[0], 1) for k, v in c.timing().items()})}", flush=True)
        except nes.NesError as e:
            print(f"dbound {dbound:g} m={m}: {e}; status {c.status} minor {c.minor} after {time.perf_counter() - t0:.2f}s", flush=True)
