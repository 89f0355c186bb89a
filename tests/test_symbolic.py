"""CPU tests of the host symbolic analysis behind nes_analyze for sparse A (csrc/sparse_symbolic.cu;
cholmod_analyze's role, sparse-cholesky.lisp:261, 509): ordering, supernodes, assembly-tree levels, the
multifrontal index maps and the subtree-to-rank mapping.  The maps are exercised by a NumPy emulation
of the device's multifrontal numeric phase (same storage, same extend-add rules), whose factor must
satisfy L L' = P M P'.  No GPU needed."""
import numpy as np
import pytest
import scipy.sparse as sp

from cholesky_is_magic_b200 import nes
from oracle import newton_solve as ons

from mf_emulation import Emu, factor_and_solve_all_ranks


def banded(rng, m, n, bw, per_col):
    rows, cols, vals = [], [], []
    for j in range(n):
        c = j * m // n if j >= m else j
        r = np.unique(np.clip(c + rng.integers(-bw, bw + 1, per_col - 1), 0, m - 1))
        r = np.union1d(r, [c])
        rows += r.tolist(); cols += [j] * len(r); vals += (1 + rng.random(len(r))).tolist()
    return sp.csc_matrix((vals, (rows, cols)), shape=(m, n))


def analyze(A, nranks=1, leaf=0):
    A = sp.csc_matrix(A)
    A.sort_indices()
    return nes.symbolic_analyze(A.indptr, A.indices, A.shape[0], A.shape[1], nranks, leaf)


def true_colcounts(pattern):
    """Boolean right-looking elimination of a symmetric pattern (dense, small m)."""
    m = pattern.shape[0]
    F = np.tril(pattern | np.eye(m, dtype=bool))
    for j in range(m):
        below = np.nonzero(F[j + 1:, j])[0] + j + 1
        if len(below):
            F[np.ix_(below, below)] |= np.tril(np.ones((len(below), len(below)), dtype=bool))
    return F.sum(axis=0), F


CASES = [("random", 1, 1, 0), ("random", 7, 12, 0), ("random", 40, 90, 0), ("random", 150, 400, 0),
         ("banded", 300, 700, 0), ("banded", 600, 1500, 32), ("banded", 900, 2000, 48)]


@pytest.mark.parametrize("kind,m,n,leaf", CASES)
def test_symbolic_structure_and_multifrontal_maps(kind, m, n, leaf):
    rng = np.random.default_rng(m)
    A = ons.random_sparse_matrix(rng, m, n, 0.05 if m <= 60 else 0.02) if kind == "random" else banded(rng, m, n, 12, 5)
    A = sp.csc_matrix(A)
    S = analyze(A, leaf=leaf)
    perm = S["perm"]
    assert sorted(perm.tolist()) == list(range(m))
    pat = ((abs(A) @ abs(A).T).toarray() != 0)
    cc, F = true_colcounts(pat[np.ix_(perm, perm)])
    assert S["anz"] == np.count_nonzero(np.tril(pat))
    assert S["aatfl"] == float((np.diff(A.indptr).astype(float) ** 2).sum())
    assert S["lnz"] == cc.sum() and S["fl"] == float((cc.astype(float) ** 2).sum())
    ns = int(S["nsuper"])
    first, nr, rowptr, rows = S["first"], S["nr"], S["rowptr"], S["rows"]
    assert first[0] == 0 and first[ns] == m and np.all(np.diff(first) >= 1) and np.all(np.diff(first) <= 128)
    for s in range(ns):
        R = rows[rowptr[s]: rowptr[s + 1]]
        assert np.array_equal(R[: first[s + 1] - first[s]], np.arange(first[s], first[s + 1]))
        assert np.all(np.diff(R) > 0)
        last = first[s + 1] - 1
        assert np.array_equal(R[first[s + 1] - first[s]:], np.nonzero(F[last + 1:, last])[0] + last + 1)
        for j in range(first[s], first[s + 1]):                      # relaxed supernodes only add zeros
            assert set(np.nonzero(F[j:, j])[0] + j) <= set(R.tolist())
        p = S["sparent"][s]
        assert (p == -1) == (nr[s] == first[s + 1] - first[s])
        if p >= 0:
            assert S["level"][p] > S["level"][s] and p > s
    lv = S["lvlptr"]
    assert lv[0] == 0 and lv[-1] == ns
    for l in range(int(S["nlevels"])):
        assert np.all(S["level"][lv[l]: lv[l + 1]] == l)
    assert np.all(S["off"] % 16 == 0) and np.all(S["ld"] % 16 == 0)
    # numeric emulation on the maps
    s_ = np.sqrt(0.1 + 10 * rng.random(n))
    M = ons.normal_matrix(A, s_)
    b = rng.random(m)
    L, x = factor_and_solve_all_ranks(S, M, b, 1)
    Mp = M[np.ix_(perm, perm)]
    assert np.linalg.norm(L @ L.T - Mp) / np.linalg.norm(Mp) <= 1e-12
    assert np.linalg.norm(M @ x - b) / np.linalg.norm(b) <= 1e-10
    # slab tables: rows of every child that land in each 64-row slab of the parent's rows below
    nc_all = np.diff(S["first"])
    for c in range(ns):
        p = S["sparent"][c]
        tb = S["tb"][S["tbptr"][c]: S["tbptr"][c + 1]]
        if p < 0 or nr[c] == nc_all[c]:
            assert len(tb) == 0
            continue
        rel = S["rel"][S["relptr"][c]: S["relptr"][c + 1]]
        nup = nr[p] - nc_all[p]
        assert len(tb) == (nup + 63) // 64 + 1 and tb[-1] == len(rel)
        for k in range(len(tb)):
            assert tb[k] == S["cut"][c] + np.searchsorted(rel[S["cut"][c]:], nc_all[p] + 64 * k)
    assert np.all(S["nb0"] % 32 == 0) and np.all(S["nb0"] >= nc_all) and np.all(S["ld"] >= S["nb0"] + nr - nc_all)


def test_nested_dissection_gives_parallel_levels_and_reuses_update_slots():
    rng = np.random.default_rng(3)
    m, n = 4000, 9000
    A = banded(rng, m, n, 20, 6)
    S_nd = analyze(A, leaf=250)
    S_rcm = analyze(A, leaf=10 ** 9)          # no dissection: profile ordering, a chain
    width_nd = np.diff(S_nd["lvlptr"]).max()
    assert np.median(np.diff(S_rcm["lvlptr"])) == 1          # a chain
    assert width_nd >= 8 and S_nd["nlevels"] < S_rcm["nlevels"] / 2
    assert S_nd["lnz"] < 6 * S_rcm["lnz"]
    total_u = sum(int(S_nd["ldu"][s]) * int(S_nd["nr"][s] - (S_nd["first"][s + 1] - S_nd["first"][s]))
                  for s in range(int(S_nd["nsuper"])))
    assert S_nd["usize"] < total_u             # slots are reused across levels


@pytest.mark.parametrize("Q", [2, 4, 8])
def test_subtree_to_rank_mapping(Q):
    rng = np.random.default_rng(5)
    m, n = 1500, 3500
    A = banded(rng, m, n, 8, 5)
    S = analyze(A, nranks=Q, leaf=40)
    owner, par = S["owner"], S["sparent"]
    ns = int(S["nsuper"])
    assert owner.min() >= -1 and owner.max() < Q
    assert set(owner[owner >= 0].tolist()) == set(range(Q))          # every rank owns something
    for s in range(ns):
        p = par[s]
        if p >= 0:
            if owner[s] == -1:
                assert owner[p] == -1                                 # the top is closed under "parent"
            else:
                assert owner[p] in (-1, owner[s])                     # subtrees are not split across ranks
    # the ordering / structure itself does not depend on the number of ranks
    S1 = analyze(A, nranks=1, leaf=40)
    for k in ("perm", "first", "rows", "level", "rel"):
        assert np.array_equal(S[k], S1[k])
    # numeric phase + solves with the ranks' phases and exchanges emulated in NumPy
    s_ = np.sqrt(0.1 + 10 * rng.random(n))
    M = ons.normal_matrix(A, s_)
    b = rng.random(m)
    L, x = factor_and_solve_all_ranks(S, M, b, Q)
    Mp = M[np.ix_(S["perm"], S["perm"])]
    assert np.linalg.norm(L @ L.T - Mp) / np.linalg.norm(Mp) <= 1e-12
    # backward-stable solve: residual small against |M| |x| (this banded M is ill-conditioned, |x| ~ 1e6 |b|)
    assert np.linalg.norm(M @ x - b) <= 1e-13 * np.linalg.norm(M) * np.linalg.norm(x)
    assert S["xu_off"][-1] == S["usize"] and len(S["xu_off"]) == Q + 1
    # load balance of the owned flops (nc * nr^2 model): no rank above 1.6x the mean
    nc = np.diff(S["first"]).astype(float)
    fl = nc * S["nr"].astype(float) ** 2
    load = np.array([fl[owner == q].sum() for q in range(Q)])
    assert load.max() <= 1.6 * load.mean()


def _edge_matrices():
    rng = np.random.default_rng(9)
    out = {}
    # one fully dense row (goes last in the ordering), the rest banded
    A = banded(rng, 500, 1100, 6, 4).tolil()
    A[137, :] = 1.0 + rng.random(1100)
    out["dense_row"] = sp.csc_matrix(A)
    # three disconnected components (independent subtrees of the elimination forest) + an isolated row
    blocks = [banded(rng, k, 2 * k + 3, 5, 4) for k in (150, 90, 260)] + [sp.csc_matrix(np.array([[2.0]]))]
    out["components"] = sp.block_diag(blocks, format="csc")
    # dense-ish: chains longer than 128 columns must be split into several supernodes
    out["dense_block"] = sp.csc_matrix(rng.random((300, 400)) * (rng.random((300, 400)) < 0.3) + np.eye(300, 400))
    return out


@pytest.mark.parametrize("name", ["dense_row", "components", "dense_block"])
def test_symbolic_edge_patterns(name):
    A = _edge_matrices()[name]
    m, n = A.shape
    S = analyze(A, leaf=64)
    perm = S["perm"]
    assert sorted(perm.tolist()) == list(range(m))
    if name == "dense_row":
        assert perm[-1] == 137                      # the dense row is eliminated last
    if name == "components":
        assert (S["sparent"] == -1).sum() >= 4      # a forest: one root per component
    if name == "dense_block":
        assert np.diff(S["first"]).max() == 128 and S["nsuper"] >= 3
    rng = np.random.default_rng(1)
    s_ = np.sqrt(0.1 + 10 * rng.random(n))
    M = ons.normal_matrix(A, s_)
    b = rng.random(m)
    L, x = factor_and_solve_all_ranks(S, M, b, 1)
    Mp = M[np.ix_(perm, perm)]
    assert np.linalg.norm(L @ L.T - Mp) / np.linalg.norm(Mp) <= 1e-12
    assert np.linalg.norm(M @ x - b) <= 1e-12 * np.linalg.norm(M) * np.linalg.norm(x)
    # the same patterns sharded over 3 ranks (odd count, forest with several roots)
    S3 = analyze(A, nranks=3, leaf=64)
    L3, x3 = factor_and_solve_all_ranks(S3, M, b, 3)
    Mp3 = M[np.ix_(S3["perm"], S3["perm"])]
    assert np.linalg.norm(L3 @ L3.T - Mp3) / np.linalg.norm(Mp3) <= 1e-12
    assert np.linalg.norm(M @ x3 - b) <= 1e-12 * np.linalg.norm(M) * np.linalg.norm(x3)


def test_symbolic_degenerate_shapes():
    # no columns at all, a single entry, an empty row: the analysis must still produce a valid structure
    for A in (sp.csc_matrix((4, 0)), sp.csc_matrix(np.array([[3.0]])),
              sp.csc_matrix(np.array([[1.0, 2.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 1.0]]))):
        S = analyze(A)
        m = A.shape[0]
        assert sorted(S["perm"].tolist()) == list(range(m))
        assert S["first"][-1] == m and S["lnz"] >= m


_THREAD_PROBE = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, {root!r})
import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, nes
import scipy.sparse as sp
sf = lpgen.sparse_lp(24000, 60000, nnz_per_col=8, bandwidth=120, seed=3)
A = sp.csc_matrix((sf.A.value, (sf.A.row, sf.A.col)), shape=(24000, 60000)); A.sort_indices()
S = nes.symbolic_analyze(A.indptr, A.indices, 24000, 60000, 1, 0)
h = hashlib.sha1()
for k in ("perm", "first", "rows", "rowptr", "level", "child", "rel", "tb", "ei", "ej", "edest", "off", "uoff"):
    h.update(np.ascontiguousarray(S[k]).tobytes())
print(h.hexdigest(), int(S["nsuper"]), int(S["nlevels"]), S["lnz"])
"""


def test_analysis_does_not_depend_on_the_number_of_host_threads():
    """Pattern, dissection branches, row structures and the assembly map run on host threads
    (NES_HOST_THREADS, read once per process), and so do the top-level searches of the dissection once
    a vertex set has 20 000 vertices: every thread count must give the same arrays."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for nth in ("1", "3", "8"):
        env = dict(os.environ, NES_HOST_THREADS=nth)
        r = subprocess.run([sys.executable, "-c", _THREAD_PROBE.format(root=root)], env=env, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1] == outs[2], outs
    assert int(outs[0].split()[2]) > 3        # a real tree, not a chain of one level


def test_elimination_tree_from_A_equals_the_tree_of_the_pattern_of_AAt(monkeypatch):
    """The analysis builds the elimination tree and the column counts from A itself (column elimination
    tree of A' and the A'A variant of the skeleton counts, O(nnz(A))) instead of walking the pattern of
    A A'.  NES_SYMBOLIC_CHECK makes it compute both ways and fail on any difference."""
    monkeypatch.setenv("NES_SYMBOLIC_CHECK", "1")
    rng = np.random.default_rng(17)
    cases = list(_edge_matrices().values())
    cases += [banded(rng, 700, 1600, 24, 5), banded(rng, 2500, 5200, 9, 4),
              sp.random(400, 900, density=0.01, random_state=3, format="csc") + sp.eye(400, 900, format="csc"),
              sp.csc_matrix(np.array([[1.0, 2.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 1.0]]))]
    for A in cases:
        for leaf in (0, 64):
            S = analyze(A, leaf=leaf)      # raises NesError if the two trees differ
            assert sorted(S["perm"].tolist()) == list(range(A.shape[0]))
