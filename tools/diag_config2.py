"""GPU PDAS on BASELINE config 2, iteration by iteration, beside the oracle's golden run."""
import json, sys, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, nes, pdas
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
from cholesky_is_magic_b200.standard_form import StandardForm
g = json.load(open("tests/golden/pdas_dense_m8192_n16384_seed0.json"))
m, n, seed = g["m"], g["n"], g["seed"]
with with_cholmod(device=0) as c:
    A = nes.Matrix.generate_dense(c, m, n, seed)
    xs, ys, zs = lpgen.aux_vectors(m, n, seed)
    b = A.sdmult(xs); cvec = A.sdmult(ys, transpose=True) + zs; A.free()
    sf = StandardForm(nvars=n, ncons=m, c=list(enumerate(cvec.tolist())), A=None, b=b, l=np.zeros(n), u=np.full(n, np.inf), initial_vars=n)
    st = pdas.make_pdas(sf, scale=True, generated_seed=seed)
    obj, gap, it = pdas.pdas(st, 300)
    print("gpu iterations", it, "obj", obj, "gap", gap, "| oracle", g["iterations"], g["dobj"], g["final_gap"])
    for i, e in enumerate(st.log):
        og = g["gaps"][i] if i < len(g["gaps"]) else None
        os_ = g["steps"][i] if i < len(g["steps"]) else None
        ob = g["branches"][i] if i < len(g["branches"]) else None
        print(i + 1, e["branch"], ob, "gap %.6e" % e["gap"], "oracle", None if og is None else "%.6e" % og, "step", e.get("step"), os_,
              "viol", ["%.2e" % v for v in e["violations"]], "pobj %.10g dobj %.10g" % (e["pobj"], e["dobj"]))
