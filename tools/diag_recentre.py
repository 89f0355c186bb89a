"""Per-iteration comparison of the GPU PDAS with the oracle on the LPs that reach the recentre branch."""
import sys, copy, warnings, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
warnings.filterwarnings("ignore")
from cholesky_is_magic_b200 import lpgen, pdas
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
from oracle import pdas as opdas
for seed, m, n in [(30, 12, 30), (32, 20, 50)]:
    sf = lpgen.dense_lp(m, n, seed); rng = np.random.default_rng(seed); sf = copy.copy(sf)
    sf.u = np.where(rng.random(n) < 0.5, 20.0, np.inf); k = rng.integers(0, n, 3)
    ost = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), sf.A_dense, sf.b, sf.l, sf.u); ost.x[k] = ost.l[k] + 1e-9
    with with_cholmod(device=0) as c:
        st = pdas.make_pdas(sf); st.x0[k] = st.l[k] + 1e-9; st.handle()
        rep_o = rep_g = False
        for it in range(70):
            go, do, so = opdas.one_pdas_iteration(ost, rep_o)
            gg, dg, sg = pdas.one_pdas_iteration(st, rep_g)
            rep_o = so is not None and so < 1e-6; rep_g = sg is not None and sg < 1e-6
            xd = np.abs(st.get("x") - ost.x).max(); zd = np.abs(st.get("z") - ost.z).max(); wd = np.abs(st.get("w") - ost.w).max(); yd = np.abs(st.get("y") - ost.y).max()
            print(seed, it, ost.log[-1]["branch"], st.log[-1]["branch"], "gap %.3e %.3e" % (go, gg), "dobj %.12g %.12g" % (do, dg), "step", so, sg, "dx %.1e dz %.1e dw %.1e dy %.1e" % (xd, zd, wd, yd), flush=True)
            if go < 1e-4 or gg < 1e-4: break
        pdas.free_pdas_A(st)
