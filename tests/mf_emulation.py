"""NumPy emulation of the device's multifrontal numeric phase and solves on the symbolic structure that
csrc/sparse_symbolic.cu produces (same storage, same extend-add rules, same phase structure as
csrc/sparse_chol.cu: own subtrees -> exchange of the subtree roots' update matrices / vectors ->
replicated top).  Test infrastructure: it validates the index maps, the pool and exchange-region layout
and the subtree-to-rank mapping on CPU, single- or multi-rank."""
import numpy as np


class Emu:
    def __init__(self, S, rank=0):
        self.S, self.rank = S, rank
        self.ns = int(S["nsuper"])
        self.nc = np.diff(S["first"])
        self.nu = S["nr"] - self.nc
        self.Lv = np.zeros(int(S["lsize"]))
        self.U = np.full(int(S["usize"]), np.nan)
        self.uvec = np.full(int(S["vptr"][-1]) if len(S["vptr"]) else 0, np.nan)

    # ---- views ------------------------------------------------------------------------------------------
    def block(self, s):
        S = self.S
        return self.Lv[S["off"][s]: S["off"][s] + S["ld"][s] * self.nc[s]].reshape(self.nc[s], S["ld"][s]).T

    def umat(self, s):
        S = self.S
        return self.U[S["uoff"][s]: S["uoff"][s] + S["ldu"][s] * self.nu[s]].reshape(self.nu[s], S["ldu"][s]).T

    def uv(self, s):
        return self.uvec[self.S["vptr"][s]: self.S["vptr"][s] + self.nu[s]]

    def children(self, s):
        return self.S["child"][self.S["childptr"][s]: self.S["childptr"][s + 1]]

    def rel(self, c):
        return self.S["rel"][self.S["relptr"][c]: self.S["relptr"][c + 1]]

    def phase(self, ph):
        """Supernodes of phase 0 (this rank's subtrees) or 1 (replicated top), level by level."""
        S = self.S
        for lvl in range(int(S["nlevels"])):
            for s in range(S["lvlptr"][lvl], S["lvlptr"][lvl + 1]):
                if (S["owner"][s] == self.rank) if ph == 0 else (S["owner"][s] < 0):
                    yield s

    # ---- numeric factorization --------------------------------------------------------------------------
    def assemble(self, M):
        self.Lv[:] = 0.0
        self.Lv[self.S["edest"]] = M[self.S["ei"], self.S["ej"]]

    def factor_phase(self, ph):
        S = self.S
        for s in self.phase(ph):
            nc, nu, nb0 = self.nc[s], self.nu[s], S["nb0"][s]
            full = self.block(s)
            D = full[:nc, :].copy()
            Bm = full[nb0:nb0 + nu, :].copy()
            Us = None
            if nu:
                Us = self.umat(s)
                Us[:nu, :] = 0.0
            for c in self.children(s):
                Uc, rl, cut = self.umat(c), self.rel(c), S["cut"][c]
                assert not np.isnan(np.tril(Uc[:self.nu[c], :self.nu[c]])).any(), (s, c)
                for j in range(self.nu[c]):
                    i = np.arange(j, self.nu[c])
                    if j < cut:
                        lo = rl[i] < nc
                        D[rl[i[lo]], rl[j]] += Uc[i[lo], j]
                        Bm[rl[i[~lo]] - nc, rl[j]] += Uc[i[~lo], j]
                    else:
                        Us[rl[i] - nc, rl[j] - nc] += Uc[i, j]
            D = np.tril(D)
            Ld = np.linalg.cholesky(D + np.tril(D, -1).T)
            full[:nc, :] = Ld
            if nu:
                Bm = np.linalg.solve(Ld, Bm.T).T
                full[nb0:nb0 + nu, :] = Bm
                Us[:nu, :nu] -= np.tril(Bm @ Bm.T)

    def exchange_regions(self, what):
        """[(owner, slice)] of the exchange regions of `what` in ('U', 'uvec')."""
        key = "xu_off" if what == "U" else "xv_off"
        off = self.S[key]
        return [(q, slice(int(off[q]), int(off[q + 1]))) for q in range(len(off) - 1)]

    # ---- solves ------------------------------------------------------------------------------------------
    def fwd_phase(self, ph, x):
        S = self.S
        for s in self.phase(ph):
            nc, nu, nb0, c0 = self.nc[s], self.nu[s], S["nb0"][s], S["first"][s]
            full = self.block(s)
            rhs = x[c0:c0 + nc].copy()
            for c in self.children(s):
                rl, cut = self.rel(c), S["cut"][c]
                rhs[rl[:cut]] -= self.uv(c)[:cut]
            y = np.linalg.solve(np.tril(full[:nc, :]), rhs)
            x[c0:c0 + nc] = y
            if nu:
                u = full[nb0:nb0 + nu, :] @ y
                for c in self.children(s):
                    rl, cut = self.rel(c), S["cut"][c]
                    u[rl[cut:] - nc] += self.uv(c)[cut:]
                self.uv(s)[:] = u

    def bwd_phase(self, ph, x):
        S = self.S
        for s in reversed(list(self.phase(ph))):
            nc, nu, nb0, c0 = self.nc[s], self.nu[s], S["nb0"][s], S["first"][s]
            full = self.block(s)
            R = S["rows"][S["rowptr"][s] + nc: S["rowptr"][s + 1]]
            rhs = x[c0:c0 + nc] - full[nb0:nb0 + nu, :].T @ x[R]
            x[c0:c0 + nc] = np.linalg.solve(np.tril(full[:nc, :]).T, rhs)

    def dense_factor(self):
        S = self.S
        m = len(S["perm"])
        L = np.zeros((m, m))
        for s in range(self.ns):
            nc, nu, nb0 = self.nc[s], self.nu[s], S["nb0"][s]
            full = self.block(s)
            R = S["rows"][S["rowptr"][s]: S["rowptr"][s + 1]]
            for cc in range(nc):
                L[R[cc:nc], S["first"][s] + cc] = full[cc:nc, cc]
                L[R[nc:], S["first"][s] + cc] = full[nb0:nb0 + nu, cc]
        return L


def factor_and_solve_all_ranks(S, M, b, nranks):
    """Run every rank of an `nranks` job in this process (exchange = array copies).  Returns the dense
    factor assembled from the owners' blocks and every rank's solution of M x = b (original numbering)."""
    emus = [Emu(S, r) for r in range(nranks)]
    perm = S["perm"]
    for e in emus:
        e.assemble(M)
        e.factor_phase(0)
    if nranks > 1:
        for q, sl in emus[0].exchange_regions("U"):
            for e in emus:
                e.U[sl] = emus[q].U[sl]
    for e in emus:
        e.factor_phase(1)
    xs = [b[perm].astype(float).copy() for _ in emus]
    for e, x in zip(emus, xs):
        e.fwd_phase(0, x)
    if nranks > 1:
        for q, sl in emus[0].exchange_regions("uvec"):
            for e in emus:
                e.uvec[sl] = emus[q].uvec[sl]
    for e, x in zip(emus, xs):
        e.fwd_phase(1, x)
        e.bwd_phase(1, x)
        e.bwd_phase(0, x)
    owner = S["owner"]
    first = S["first"]
    # all-reduce of the masked pieces (rank 0 keeps the top)
    xsum = np.zeros(len(perm))
    for r, x in enumerate(xs):
        for s in range(int(S["nsuper"])):
            if owner[s] == r or (owner[s] < 0 and r == 0):
                xsum[first[s]: first[s + 1]] += x[first[s]: first[s + 1]]
    out = np.empty(len(perm))
    out[perm] = xsum
    # dense factor from the owners' blocks (top from rank 0)
    Lv = emus[0].Lv.copy()
    for s in range(int(S["nsuper"])):
        o = owner[s]
        if o > 0:
            sl = slice(int(S["off"][s]), int(S["off"][s + 1]))
            Lv[sl] = emus[o].Lv[sl]
    e = Emu(S, 0)
    e.Lv = Lv
    return e.dense_factor(), out
