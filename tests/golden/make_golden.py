"""Regenerates the golden fixtures from the oracle (the reference ships none and cannot run here;
see oracle/__init__.py).  Usage: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from cholesky_is_magic_b200 import lpgen  # noqa: E402
from oracle import newton_solve as ons  # noqa: E402
from oracle import pdas as opdas  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pdas_case(m, n, seed, suffix=""):
    sf = lpgen.dense_lp(m, n, seed)
    st = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), sf.A_dense, sf.b, sf.l, sf.u)
    obj, gap, iters = opdas.pdas(st, 500)
    gaps = [e["gap"] for e in st.log]
    out = {
        "m": m, "n": n, "seed": seed, "iterations": iters, "dobj": obj, "final_gap": gap,
        "pobj_last": st.log[-1]["pobj"],
        "branches": [e["branch"] for e in st.log],
        "gaps": gaps,
        "steps": [e.get("step") for e in st.log],
        # margin of the stopping rule (gap < 1e-4): how far the last two gaps are from the threshold
        "stop_margin": {"previous_gap": gaps[-2], "last_gap": gaps[-1]},
        "x_head": st.x[:8].tolist(), "y_head": st.y[:8].tolist(),
    }
    json.dump(out, open(os.path.join(HERE, f"pdas_dense_m{m}_n{n}_seed{seed}{suffix}.json"), "w"), indent=1)
    print(m, n, seed, iters, obj, gaps[-2], gaps[-1])


def kkt_case(m, n, seed):
    rng = np.random.default_rng(seed)
    l, u, w, z, A, e, f, g, h = ons.random_dense_case(rng, m, n)
    dw, dx, dy, dz, inter = ons.solve_kkt_newton(l, u, w, z, A, e, f, g, h, return_intermediates=True)
    np.savez_compressed(os.path.join(HERE, f"kkt_dense_m{m}_n{n}_seed{seed}.npz"),
                        l=l, u=u, w=w, z=z, A=A, e=e, f=f, g=g, h=h, dw=dw, dx=dx, dy=dy, dz=dz,
                        theta=inter["theta"], rhs=inter["rhs"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "config2":
        # BASELINE config 2 on the host: ~100 oracle iterations of 2.4e12 flops each (a quarter of an hour on
        # 8 cores); run once, the fixture pins the GPU's iteration count / objective at full size
        # a second argument names a variant run (e.g. under another OPENBLAS_NUM_THREADS: a different
        # summation order in dgemm / dpotrf) whose fixture shows how far the oracle moves against itself
        pdas_case(8192, 16384, 0, suffix=("_" + sys.argv[2]) if len(sys.argv) > 2 else "")
        sys.exit(0)
    pdas_case(20, 50, 0)
    pdas_case(200, 500, 0)
    kkt_case(12, 30, 5)
    kkt_case(130, 300, 6)
