"""Small end-to-end pass over the dense (two-stream look-ahead), KKT, sparse and PDAS paths in a few
seconds: a quick check after a kernel change, sized so that it would also run under compute-sanitizer
(closed on this pool: the tool answers that it stays closed, so bounds and races are checked by the
parity tests instead)."""
import sys, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, nes, newton_solve, pdas
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
from oracle import newton_solve as ons

with with_cholmod(device=0) as c:
    rng = np.random.default_rng(0)
    # dense: formation + look-ahead Cholesky (m > 1024 takes the two-stream path) + solves
    m, n = 1280, 1536
    A = rng.random((m, n)) + np.eye(m, n)
    s = np.sqrt(0.1 + 10 * rng.random(n))
    Ad = nes.Matrix.from_dense(c, A); Ad.scale(s)
    L = nes.Factor(c, Ad); assert L.factorize(Ad)
    x = L.solve(rng.random(m)); L.free(); Ad.free()
    print("dense ok", flush=True)
    # KKT Newton step (GEMVs + elementwise kernels)
    l, u, w, z, A, e, f, g, h = ons.random_dense_case(rng, 96, 224)
    Ad = nes.Matrix.from_dense(c, A)
    newton_solve.solve_kkt_newton(l, u, w, z, Ad, e, f, g, h); Ad.free()
    print("kkt ok", flush=True)
    # sparse: symbolic + multifrontal numeric + solves
    sf = lpgen.sparse_lp(3000, 7000, nnz_per_col=6, bandwidth=60, seed=1)
    As = nes.Matrix.from_triplets(c, sf.A.row, sf.A.col, sf.A.value, 3000, 7000)
    As.scale(np.sqrt(0.1 + 10 * rng.random(7000)))
    Ls = nes.Factor(c, As); assert Ls.factorize(As)
    Ls.solve(rng.random(3000)); Ls.free(); As.free()
    print("sparse ok", flush=True)
    # a short PDAS run (device-resident state, reductions)
    sfd = lpgen.dense_lp(48, 120, 2)
    pdas.pdas(pdas.make_pdas(sfd), 8, native_loop=True)
    print("pdas ok", flush=True)
