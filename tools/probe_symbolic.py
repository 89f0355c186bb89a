import sys, time, numpy as np
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, nes
import scipy.sparse as sp
m, n = 100000, 250000
sf = lpgen.sparse_lp(m, n, 10, bandwidth=400, seed=0)
A = sp.csc_matrix((sf.A.value, (sf.A.row, sf.A.col)), shape=(m, n))
A.sort_indices()
t=time.time()
s = nes.symbolic_analyze(A.indptr, A.indices, m, n)
print("analyze", time.time()-t, "s; nsuper", s["nsuper"], "nlevels", s["nlevels"], "lnz %.3g fl %.3g" % (s["lnz"], s["fl"]))
first, nr, lvlptr = s["first"], s["nr"], s["lvlptr"]
nc = np.diff(first)
for l in range(int(s["nlevels"])):
    a, b = lvlptr[l], lvlptr[l+1]
    ncs = nc[a:b]; nrs = nr[a:b]
    nu = nrs - ncs
    fl = (ncs.astype(float)**3/3 + nu.astype(float)*ncs**2 + nu.astype(float)**2*ncs).sum()
    print(f"level {l:2d}: {b-a:4d} supernodes  nc {ncs.min():3d}..{ncs.max():3d}  rows below {nu.min():5d}..{nu.max():5d}  flops {fl:.3g}")
import hashlib
print("perm sha1", hashlib.sha1(s["perm"].tobytes()).hexdigest(), "first sha1", hashlib.sha1(s["first"].tobytes()).hexdigest())
print("rows sha1", hashlib.sha1(s["rows"].tobytes()).hexdigest(), "edest sha1", hashlib.sha1(s["edest"].tobytes()).hexdigest())
