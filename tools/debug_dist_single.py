"""The distributed schedule on ONE GPU (NES_FORCE_DIST): factor with the 128x128 update kernel and with the
half-tile kernel, compare the two factors tile by tile."""
import os, sys, numpy as np
os.environ["NES_FORCE_DIST"] = "1"
os.environ["NES_DIST_NBO"] = sys.argv[3] if len(sys.argv) > 3 else "512"
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
m = int(sys.argv[1]); n = int(sys.argv[2])
mode = sys.argv[4] if len(sys.argv) > 4 else "1"
with with_cholmod(device=0, timing=True) as c:
    A = nes.Matrix.generate_dense(c, m, n, 0)
    A.scale(np.sqrt(0.1 + 10 * np.random.default_rng(0).random(n)))
    L = nes.Factor(c, A)
    os.environ["NES_UPDATE_KERNEL"] = "128"
    assert L.factorize(A); ra = L.residual(A); La = L.to_dense()
    assert L.factorize(A); La2 = L.to_dense(); print("128 vs 128 identical:", np.array_equal(La, La2)); del La2
    os.environ["NES_DIST_KERNEL_DEBUG"] = mode   # 1: half-tile kernel on the trailing updates, 2: on the chain, 3: both
    assert L.factorize(A); rb = L.residual(A); Lb = L.to_dense()
    print(f"m={m} residual 128: {ra:.2e}  64(mode {mode}): {rb:.2e}", flush=True)
    tm = (m + 127) // 128
    first = None
    for bj in range(tm):
        d = np.abs(La[:, bj * 128:(bj + 1) * 128] - Lb[:, bj * 128:(bj + 1) * 128])
        scale = np.abs(La[:, bj * 128:(bj + 1) * 128]).max()
        bad = [(bi, float(d[bi * 128:(bi + 1) * 128].max())) for bi in range(bj, tm) if d[bi * 128:(bi + 1) * 128].max() > 1e-11 * scale]
        if bad:
            print(f"tile column {bj} (block column {bj * 128 // int(os.environ['NES_DIST_NBO'])}): {len(bad)} of {tm - bj} tile rows differ; first {bad[:6]} scale {scale:.3g}", flush=True)
            if first is None:
                first = bj
                bi = bad[0][0]
                dd = d[bi * 128:(bi + 1) * 128]
                rows, cols = np.nonzero(dd > 1e-11 * scale)
                print("   first bad tile", bi, bj, "rows", rows.min(), rows.max(), "cols", cols.min(), cols.max(), "count", len(rows))
            if bj > first + 12:
                break
    L.free(); A.free()
