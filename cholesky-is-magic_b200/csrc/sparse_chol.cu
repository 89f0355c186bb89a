// K3/K4/K5 for a sparse (CSC) constraint matrix: supernodal Cholesky of A diag(theta) A'.
//
// Replaces what sparse-newton-solve.lisp / affine-scaling.lisp get from CHOLMOD:
//   cholmod_analyze    (sparse-cholesky.lisp:509, affine-scaling.lisp:270-271)  -> sparse_analyze  (HOST)
//   cholmod_factorize  (sparse-cholesky.lisp:512, 543)                          -> sparse_factorize (GPU)
//   cholmod_solve/2    (sparse-cholesky.lisp:515, 546)                          -> sparse_solve_inplace
// and the counters printed by the reference (anz, aatfl, lnz, fl; affine-scaling.lisp:273-279).
//
// Symbolic phase (host, once per pattern -- the pattern of A diag(theta) A' never changes):
//   pattern of tril(A A'), ordering (dense rows last + reverse Cuthill-McKee; CHOLMOD would use AMD --
//   the factor and the solution do not depend on it beyond rounding, only lnz/fl do), elimination
//   tree and column structures, fundamental supernodes merged along etree chains with relaxed
//   amalgamation (<=128 columns, <=15% explicit zeros), per-supernode row lists, relative-index maps for
//   the updates, and the destination of every entry of tril(A A') in the supernodal storage.
// Numeric phase (device): owner-computes assembly (one thread per entry of tril(M): merge-join of two
//   rows of A, so no atomics and bitwise reproducible), then supernodes in ascending (= topological)
//   order: diagonal block (potrf_block_smem), TRSM of the rows below, and the outer-product update
//   scattered into the ancestors through the relative maps.
// Storage: supernode s is a dense nr x nc column-major block (ld = nr) holding its columns of L.
#include <algorithm>
#include <numeric>
#include <queue>

#include "nes_internal.h"
#include "potrf_block.cuh"

namespace nes {

struct SparseFactor {
    int m = 0;
    int nsuper = 0;
    long long lsize = 0;  // doubles in the supernodal storage
    // host copies needed to drive the launches
    std::vector<int> first;       // nsuper+1
    std::vector<int> nr;          // rows per supernode
    std::vector<long long> off;   // nsuper+1
    std::vector<int> rowptr;      // nsuper+1 into rows
    std::vector<int> segptr;      // nsuper+1 into segments
    int nseg = 0;
    long long anz = 0;
    // device
    double* d_L = nullptr;
    double* d_dinv = nullptr;
    int* d_rows = nullptr;        // concatenated supernode row lists (permuted indices)
    int* d_seg_t_off = nullptr;   // per segment: nothing but packed params below
    long long* d_seg_toff = nullptr;  // offset of the target supernode block
    int* d_seg_tnr = nullptr;     // ld of the target block
    int* d_seg_tcol0 = nullptr;   // first column of the target supernode
    int* d_seg_j0 = nullptr;      // below-row range [j0, j1) of s whose rows are columns of the target
    int* d_seg_j1 = nullptr;
    int* d_seg_relptr = nullptr;  // offset into d_rel of rel[i - j0], i >= j0
    int* d_rel = nullptr;
    int* d_perm = nullptr;        // perm[new] = old
    // assembly: entry e of tril(P M P'): rows oi, oj of A (original numbering), destination in d_L
    int* d_ei = nullptr;
    int* d_ej = nullptr;
    long long* d_edest = nullptr;
    double* d_x = nullptr;        // permuted right-hand side / solution workspace
    int* d_info = nullptr;
    // dataflow solve (one cooperative launch per sweep)
    int* d_first = nullptr;       // nsuper+1
    int* d_nr = nullptr;          // nsuper
    long long* d_off = nullptr;   // nsuper+1
    int* d_rowptr = nullptr;      // nsuper+1
    long long* d_woff = nullptr;  // nsuper+1: offsets of the inverted diagonal blocks (nc x nc each)
    double* d_W = nullptr;
    int* d_inptr = nullptr;       // nsuper+1: incoming segments of each target supernode, ascending source
    int* d_in_s = nullptr;
    int* d_in_j0 = nullptr;
    int* d_in_j1 = nullptr;
    int* d_segptr = nullptr;      // nsuper+1 (device copy of segptr)
    int* d_seg_tid = nullptr;     // target supernode id of each outgoing segment
    int* d_flags = nullptr;
    int flag_epoch = 0;
    long long wsize = 0;
};

// ------------------------------------------------------------------------------------------------
// host: symbolic analysis
// ------------------------------------------------------------------------------------------------
static void csc_to_csr(int m, int n, const std::vector<int>& cp, const std::vector<int>& ri,
                       std::vector<int>& rp, std::vector<int>& cj) {
    rp.assign(m + 1, 0);
    cj.resize(ri.size());
    for (int r : ri) rp[r + 1]++;
    for (int i = 0; i < m; ++i) rp[i + 1] += rp[i];
    std::vector<int> next(rp.begin(), rp.end() - 1);
    for (int j = 0; j < n; ++j)
        for (int k = cp[j]; k < cp[j + 1]; ++k) cj[next[ri[k]]++] = j;
}

// full symmetric adjacency of A A' (without the diagonal), sorted per row
static void aat_pattern(int m, const std::vector<int>& cp, const std::vector<int>& ri,
                        const std::vector<int>& rp, const std::vector<int>& cj,
                        std::vector<int>& ap, std::vector<int>& ai, double& aatfl) {
    ap.assign(m + 1, 0);
    ai.clear();
    std::vector<int> mark(m, -1);
    aatfl = 0;
    for (size_t j = 0; j + 1 < cp.size(); ++j) {
        const double c = cp[j + 1] - cp[j];
        aatfl += c * c;
    }
    for (int i = 0; i < m; ++i) {
        const size_t start = ai.size();
        mark[i] = i;
        for (int q = rp[i]; q < rp[i + 1]; ++q) {
            const int k = cj[q];
            for (int t = cp[k]; t < cp[k + 1]; ++t) {
                const int r = ri[t];
                if (mark[r] != i) {
                    mark[r] = i;
                    ai.push_back(r);
                }
            }
        }
        std::sort(ai.begin() + start, ai.end());
        ap[i + 1] = (int)ai.size();
    }
}

// perm[new] = old.  Rows whose degree exceeds 10 sqrt(m) (at least 16) go last; the rest is ordered by
// reverse Cuthill-McKee, component by component, starting from a minimum-degree vertex.
static void order_rcm(int m, const std::vector<int>& ap, const std::vector<int>& ai,
                      std::vector<int>& perm) {
    perm.clear();
    perm.reserve(m);
    const double thresh = std::max(16.0, 10.0 * std::sqrt((double)m));
    std::vector<char> dense(m, 0), seen(m, 0);
    std::vector<int> deg(m);
    for (int i = 0; i < m; ++i) {
        deg[i] = ap[i + 1] - ap[i];
        if (deg[i] > thresh) dense[i] = 1;
    }
    std::vector<int> order;
    order.reserve(m);
    std::vector<int> byDeg(m);
    std::iota(byDeg.begin(), byDeg.end(), 0);
    std::stable_sort(byDeg.begin(), byDeg.end(), [&](int a, int b) { return deg[a] < deg[b]; });
    std::vector<int> nbr;
    for (int s : byDeg) {
        if (seen[s] || dense[s]) continue;
        size_t head = order.size();
        order.push_back(s);
        seen[s] = 1;
        while (head < order.size()) {
            const int v = order[head++];
            nbr.clear();
            for (int q = ap[v]; q < ap[v + 1]; ++q) {
                const int u = ai[q];
                if (!seen[u] && !dense[u]) {
                    seen[u] = 1;
                    nbr.push_back(u);
                }
            }
            std::sort(nbr.begin(), nbr.end(), [&](int a, int b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
            order.insert(order.end(), nbr.begin(), nbr.end());
        }
    }
    for (auto it = order.rbegin(); it != order.rend(); ++it) perm.push_back(*it);
    for (int i = 0; i < m; ++i)
        if (dense[i]) perm.push_back(i);
}

int sparse_analyze(nes_ctx* c, nes_matrix* A, nes_factor* L) {
    const MatrixBase* b = A->base;
    const int m = (int)b->m, n = (int)b->n;
    std::vector<int> rp, cj;
    csc_to_csr(m, n, b->h_colptr, b->h_rowidx, rp, cj);
    std::vector<int> ap, ai;
    double aatfl = 0;
    aat_pattern(m, b->h_colptr, b->h_rowidx, rp, cj, ap, ai, aatfl);
    std::vector<int> perm, iperm(m);
    order_rcm(m, ap, ai, perm);
    for (int i = 0; i < m; ++i) iperm[perm[i]] = i;

    // ---- column structures of L (permuted), elimination tree --------------------------------------
    // below-diagonal pattern of column j of P M P': { iperm[r] : r adjacent to perm[j], iperm[r] > j };
    // struct(L_j) = that pattern united with struct(L_child) \ {j} over the etree children of j.
    std::vector<std::vector<int>> lstruct(m);  // below-diagonal rows of L(:, j), sorted
    std::vector<int> parent(m, -1), colcount(m);
    std::vector<std::vector<int>> children(m);
    std::vector<int> mark(m, -1);
    long long anz = m;
    {
        std::vector<int> tmp;
        for (int j = 0; j < m; ++j) {
            tmp.clear();
            mark[j] = j;
            const int oj = perm[j];
            for (int q = ap[oj]; q < ap[oj + 1]; ++q) {
                const int i = iperm[ai[q]];
                if (i > j && mark[i] != j) {
                    mark[i] = j;
                    tmp.push_back(i);
                }
            }
            anz += (long long)tmp.size();
            for (int ch : children[j])
                for (int i : lstruct[ch])
                    if (i > j && mark[i] != j) {
                        mark[i] = j;
                        tmp.push_back(i);
                    }
            std::sort(tmp.begin(), tmp.end());
            lstruct[j] = tmp;
            colcount[j] = (int)tmp.size() + 1;
            if (!tmp.empty()) {
                parent[j] = tmp[0];
                children[tmp[0]].push_back(j);
            }
        }
    }
    double lnz = 0, fl = 0;
    for (int j = 0; j < m; ++j) {
        lnz += colcount[j];
        fl += (double)colcount[j] * colcount[j];
    }

    // ---- supernodes: chains parent(j) = j+1 with nested structure, relaxed amalgamation ---------
    SparseFactor* sf = new SparseFactor();
    sf->m = m;
    std::vector<int> first;
    first.push_back(0);
    {
        int f = 0;
        long long zeros = 0;  // explicit zeros if [f..j] is one supernode with the row set of column j
        for (int j = 0; j + 1 <= m; ++j) {
            bool merge = false;
            if (j + 1 < m && parent[j] == j + 1) {
                const int width = j + 1 - f + 1;  // width after the merge
                // rows of the merged supernode = own columns + below-structure of column j+1
                const long long nrows = width + (long long)lstruct[j + 1].size();
                long long z = 0;
                for (int q = f; q <= j + 1; ++q) z += (nrows - (q - f)) - colcount[q];
                const long long total = nrows * width - (long long)width * (width - 1) / 2;
                if (width <= CH_NB && (width <= 4 || z <= 0.15 * total)) {
                    merge = true;
                    zeros = z;
                }
            }
            if (!merge) {
                first.push_back(j + 1);
                f = j + 1;
                zeros = 0;
            }
        }
        (void)zeros;
    }
    const int nsuper = (int)first.size() - 1;
    sf->nsuper = nsuper;
    sf->first = first;
    std::vector<int> col2sn(m);
    sf->nr.resize(nsuper);
    sf->off.assign(nsuper + 1, 0);
    sf->rowptr.assign(nsuper + 1, 0);
    std::vector<int> rows;
    for (int s = 0; s < nsuper; ++s) {
        const int f = first[s], l = first[s + 1] - 1;
        for (int j = f; j <= l; ++j) {
            col2sn[j] = s;
            rows.push_back(j);
        }
        for (int i : lstruct[l]) rows.push_back(i);
        sf->nr[s] = (l - f + 1) + (int)lstruct[l].size();
        sf->rowptr[s + 1] = (int)rows.size();
        sf->off[s + 1] = sf->off[s] + (long long)sf->nr[s] * (l - f + 1);
    }
    sf->lsize = sf->off[nsuper];

    // ---- update segments + relative maps -------------------------------------------------------
    std::vector<long long> seg_toff;
    std::vector<int> seg_tnr, seg_tcol0, seg_j0, seg_j1, seg_relptr, rel, seg_tid, seg_src;
    sf->segptr.assign(nsuper + 1, 0);
    for (int s = 0; s < nsuper; ++s) {
        const int nc = first[s + 1] - first[s];
        const int* R = rows.data() + sf->rowptr[s];
        const int rb = sf->nr[s] - nc;
        int j = 0;
        while (j < rb) {
            const int t = col2sn[R[nc + j]];
            int j1 = j;
            while (j1 < rb && col2sn[R[nc + j1]] == t) ++j1;
            seg_toff.push_back(sf->off[t]);
            seg_tid.push_back(t);
            seg_src.push_back(s);
            seg_tnr.push_back(sf->nr[t]);
            seg_tcol0.push_back(first[t]);
            seg_j0.push_back(j);
            seg_j1.push_back(j1);
            seg_relptr.push_back((int)rel.size());
            const int* Rt = rows.data() + sf->rowptr[t];
            const int nrt = sf->nr[t];
            int p = 0;
            for (int i = j; i < rb; ++i) {
                while (p < nrt && Rt[p] < R[nc + i]) ++p;
                if (p >= nrt || Rt[p] != R[nc + i]) {
                    delete sf;
                    return fail(c, NES_ERR_INVALID, "symbolic analysis: row %d of supernode %d missing in ancestor %d",
                                R[nc + i], s, t);
                }
                rel.push_back(p);
            }
            j = j1;
        }
        sf->segptr[s + 1] = (int)seg_toff.size();
    }
    sf->nseg = (int)seg_toff.size();
    // incoming segments per target (sources ascending because segments are generated in source order)
    std::vector<int> inptr(nsuper + 1, 0), in_s(sf->nseg), in_j0(sf->nseg), in_j1(sf->nseg);
    for (int q = 0; q < sf->nseg; ++q) inptr[seg_tid[q] + 1]++;
    for (int t = 0; t < nsuper; ++t) inptr[t + 1] += inptr[t];
    {
        std::vector<int> next(inptr.begin(), inptr.end() - 1);
        for (int q = 0; q < sf->nseg; ++q) {
            const int p = next[seg_tid[q]]++;
            in_s[p] = seg_src[q];
            in_j0[p] = seg_j0[q];
            in_j1[p] = seg_j1[q];
        }
    }
    std::vector<long long> woff(nsuper + 1, 0);
    for (int t = 0; t < nsuper; ++t) {
        const long long nc = first[t + 1] - first[t];
        woff[t + 1] = woff[t] + nc * nc;
    }
    sf->wsize = woff[nsuper];

    // ---- assembly map: every entry (i >= j) of tril(P M P') -> position in the supernodal storage
    std::vector<int> ei, ej;
    std::vector<long long> edest;
    ei.reserve(anz);
    ej.reserve(anz);
    edest.reserve(anz);
    {
        std::vector<int> col;
        for (int j = 0; j < m; ++j) {
            const int s = col2sn[j];
            const int* R = rows.data() + sf->rowptr[s];
            const int nrs = sf->nr[s];
            const long long base = sf->off[s] + (long long)(j - first[s]) * nrs;
            col.clear();
            col.push_back(j);
            const int oj = perm[j];
            for (int q = ap[oj]; q < ap[oj + 1]; ++q) {
                const int i = iperm[ai[q]];
                if (i > j) col.push_back(i);
            }
            std::sort(col.begin(), col.end());
            int p = 0;
            for (int i : col) {
                while (p < nrs && R[p] < i) ++p;
                ei.push_back(perm[i]);
                ej.push_back(oj);
                edest.push_back(base + p);
            }
        }
    }
    sf->anz = (long long)ei.size();

    // ---- upload -------------------------------------------------------------------------------
    auto up_i = [&](int** dst, const std::vector<int>& v) {
        *dst = static_cast<int*>(dev_alloc(c, (v.size() + 1) * sizeof(int)));
        return *dst && upload(c, *dst, v.data(), v.size() * sizeof(int)) == 0;
    };
    auto up_l = [&](long long** dst, const std::vector<long long>& v) {
        *dst = static_cast<long long*>(dev_alloc(c, (v.size() + 1) * sizeof(long long)));
        return *dst && upload(c, *dst, v.data(), v.size() * sizeof(long long)) == 0;
    };
    L->sparse = sf;
    bool ok = up_i(&sf->d_rows, rows) && up_l(&sf->d_seg_toff, seg_toff) && up_i(&sf->d_seg_tnr, seg_tnr) &&
              up_i(&sf->d_seg_tcol0, seg_tcol0) && up_i(&sf->d_seg_j0, seg_j0) && up_i(&sf->d_seg_j1, seg_j1) &&
              up_i(&sf->d_seg_relptr, seg_relptr) && up_i(&sf->d_rel, rel) && up_i(&sf->d_perm, perm) &&
              up_i(&sf->d_ei, ei) && up_i(&sf->d_ej, ej) && up_l(&sf->d_edest, edest) &&
              up_i(&sf->d_first, first) && up_i(&sf->d_nr, sf->nr) && up_l(&sf->d_off, sf->off) &&
              up_i(&sf->d_rowptr, sf->rowptr) && up_l(&sf->d_woff, woff) && up_i(&sf->d_inptr, inptr) &&
              up_i(&sf->d_in_s, in_s) && up_i(&sf->d_in_j0, in_j0) && up_i(&sf->d_in_j1, in_j1) &&
              up_i(&sf->d_segptr, sf->segptr) && up_i(&sf->d_seg_tid, seg_tid);
    if (ok) {
        sf->d_L = static_cast<double*>(dev_alloc(c, (size_t)(sf->lsize + 16) * sizeof(double)));
        sf->d_dinv = static_cast<double*>(dev_alloc(c, (size_t)(m + 16) * sizeof(double)));
        sf->d_x = static_cast<double*>(dev_alloc(c, (size_t)(m + 16) * sizeof(double)));
        sf->d_info = static_cast<int*>(dev_alloc(c, 4 * sizeof(int)));
        sf->d_W = static_cast<double*>(dev_alloc(c, (size_t)(sf->wsize + 16) * sizeof(double)));
        sf->d_flags = static_cast<int*>(dev_alloc(c, (size_t)(nsuper + 1) * sizeof(int)));
        L->d_rhs = static_cast<double*>(dev_alloc(c, (size_t)(m + 16) * sizeof(double)));
        ok = sf->d_L && sf->d_dinv && sf->d_x && sf->d_info && L->d_rhs && sf->d_W && sf->d_flags;
        if (ok) cudaMemsetAsync(sf->d_flags, 0, (size_t)(nsuper + 1) * sizeof(int), c->stream);
    }
    if (!ok) return c->status < 0 ? c->status : NES_ERR_OUT_OF_MEMORY;
    c->anz = (double)sf->anz;
    c->aatfl = aatfl;
    c->lnz = lnz;
    c->fl = fl;
    c->status = 0;
    return 0;
}

void sparse_free(nes_ctx* c, nes_factor* L) {
    SparseFactor* sf = L->sparse;
    if (!sf) return;
    void* ptrs[] = {sf->d_L, sf->d_dinv, sf->d_rows, sf->d_seg_toff, sf->d_seg_tnr, sf->d_seg_tcol0,
                    sf->d_seg_j0, sf->d_seg_j1, sf->d_seg_relptr, sf->d_rel, sf->d_perm, sf->d_ei,
                    sf->d_ej, sf->d_edest, sf->d_x, sf->d_info, sf->d_first, sf->d_nr, sf->d_off, sf->d_rowptr,
                    sf->d_woff, sf->d_W, sf->d_inptr, sf->d_in_s, sf->d_in_j0, sf->d_in_j1, sf->d_segptr,
                    sf->d_seg_tid, sf->d_flags};
    for (void* p : ptrs) dev_free(c, p);
    delete sf;
    L->sparse = nullptr;
}

// ------------------------------------------------------------------------------------------------
// device: numeric factorization
// ------------------------------------------------------------------------------------------------
// M(i,j) = sum_k A(i,k) theta_k A(j,k): merge-join of rows oi and oj of A (CSR, sorted columns).
__global__ void sparse_assemble_kernel(long long anz, const int* __restrict__ ei, const int* __restrict__ ej,
                                       const long long* __restrict__ edest, const int* __restrict__ rowptr,
                                       const int* __restrict__ colidx, const double* __restrict__ val,
                                       const double* __restrict__ theta, double* __restrict__ Lv) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= anz) return;
    const int i = ei[e], j = ej[e];
    int p = rowptr[i], q = rowptr[j];
    const int pe = rowptr[i + 1], qe = rowptr[j + 1];
    double acc = 0.0;
    while (p < pe && q < qe) {
        const int cp = colidx[p], cq = colidx[q];
        if (cp == cq) {
            const double t = theta ? theta[cp] : 1.0;
            acc = fma(val[p] * t, val[q], acc);
            ++p;
            ++q;
        } else if (cp < cq) {
            ++p;
        } else {
            ++q;
        }
    }
    Lv[edest[e]] = acc;
}

constexpr int SN_DIAG_SMEM = (CH_NB * CH_P + CH_NB) * 8;

__global__ void __launch_bounds__(256)
snode_potrf_kernel(double* __restrict__ Lv, long long off, int nr, int nc, int col0,
                   double* __restrict__ dinv_out, double dbound, int* __restrict__ info) {
    extern __shared__ __align__(128) double S[];
    double* dinv = S + CH_NB * CH_P;
    const int tid = threadIdx.x;
    double* blk = Lv + off;
    for (int idx = tid; idx < nc * nc; idx += 256) {
        const int cc = idx / nc, r = idx - cc * nc;
        if (r >= cc) S[r + cc * CH_P] = blk[r + (long long)cc * nr];
    }
    __syncthreads();
    potrf_block_smem(S, dinv, nc, dbound, info, col0);
    for (int idx = tid; idx < nc * nc; idx += 256) {
        const int cc = idx / nc, r = idx - cc * nc;
        if (r >= cc) blk[r + (long long)cc * nr] = S[r + cc * CH_P];
    }
    if (tid < nc) dinv_out[col0 + tid] = dinv[tid];
}

// X L' = B for the rows below the diagonal block of a supernode (nc <= 128 columns).
constexpr int SN_TR_ROWS = 64;
constexpr int SN_TR_SMEM = (CH_NB * CH_NB + SN_TR_ROWS * CH_NB + CH_NB) * 8;

__global__ void __launch_bounds__(256)
snode_trsm_kernel(double* __restrict__ Lv, long long off, int nr, int nc, int col0,
                  const double* __restrict__ dinv_g) {
    extern __shared__ double sm[];
    double* Ls = sm;                        // Ls[c + p*128] = L[c][p]
    double* Xs = Ls + CH_NB * CH_NB;        // Xs[p*64 + row]
    double* dv = Xs + SN_TR_ROWS * CH_NB;
    const int tid = threadIdx.x;
    const int row0 = nc + blockIdx.x * SN_TR_ROWS;
    const int nrows = min(SN_TR_ROWS, nr - row0);
    double* blk = Lv + off;
    for (int idx = tid; idx < nc * CH_NB; idx += 256) {
        const int p = idx >> 7, cc = idx & 127;
        Ls[idx] = (cc >= p && cc < nc) ? blk[cc + (long long)p * nr] : 0.0;
    }
    if (tid < CH_NB) dv[tid] = (tid < nc) ? dinv_g[col0 + tid] : 1.0;
    for (int idx = tid; idx < SN_TR_ROWS * nc; idx += 256) {
        const int p = idx >> 6, rr = idx & 63;
        if (rr < nrows) Xs[idx] = blk[row0 + rr + (long long)p * nr];
    }
    __syncthreads();
    trsm_slab_smem(Ls, Xs, dv, nc, nrows);
    __syncthreads();
    for (int idx = tid; idx < SN_TR_ROWS * nc; idx += 256) {
        const int p = idx >> 6, rr = idx & 63;
        if (rr < nrows) blk[row0 + rr + (long long)p * nr] = Xs[idx];
    }
}

// Outer-product update of the ancestors: for segment q of supernode s (rows j0..j1 of the below part
// are columns of ancestor t), U(i, j) = sum_c B(i,c) B(j,c) for i >= j, j in [j0, j1), is subtracted
// from L_t at (rel[i - j0], column R[j] - first[t]).  grid = (row tiles of 64, segments of s).
// Each CTA stages the 64-row slab and the segment's rows of B in shared memory (k <= 128).
constexpr int SN_UP_ROWS = 64;
constexpr int SN_UP_SMEM = (SN_UP_ROWS * (CH_NB + 1) + CH_NB * (CH_NB + 1)) * 8;

__global__ void __launch_bounds__(256)
snode_update_kernel(double* __restrict__ Lv, long long off, int nr, int nc, const int* __restrict__ rows,
                    int seg0, const long long* __restrict__ seg_toff, const int* __restrict__ seg_tnr,
                    const int* __restrict__ seg_tcol0, const int* __restrict__ seg_j0,
                    const int* __restrict__ seg_j1, const int* __restrict__ seg_relptr,
                    const int* __restrict__ rel) {
    extern __shared__ double sm[];
    const int P = CH_NB + 1;
    double* Bi = sm;                    // Bi[r * P + c], 64 rows of the slab
    double* Bj = Bi + SN_UP_ROWS * P;   // Bj[j * P + c], the segment's rows (<= 128)
    const int seg = seg0 + blockIdx.y;
    const int j0 = seg_j0[seg], j1 = seg_j1[seg];
    const int rb = nr - nc;
    const int i0 = j0 + blockIdx.x * SN_UP_ROWS;
    if (i0 >= rb) return;
    const int ni = min(SN_UP_ROWS, rb - i0);
    const int nj = j1 - j0;
    const int tid = threadIdx.x;
    const double* B = Lv + off + nc;  // B(i, c) = B[i + c*nr]
    for (int idx = tid; idx < SN_UP_ROWS * nc; idx += 256) {
        const int cc = idx >> 6, r = idx & 63;
        Bi[r * P + cc] = (r < ni) ? B[i0 + r + (long long)cc * nr] : 0.0;
    }
    for (int idx = tid; idx < nj * nc; idx += 256) {
        const int cc = idx / nj, j = idx - cc * nj;
        Bj[j * P + cc] = B[j0 + j + (long long)cc * nr];
    }
    __syncthreads();
    // thread (ti, tj): rows ti + 16a (a < 4), cols tj + 16b (b < 8)
    const int ti = tid & 15, tj = tid >> 4;
    double acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) acc[a][b2] = 0.0;
    const int nb = (nj + 15) >> 4;
    for (int cc = 0; cc < nc; ++cc) {
        double xi[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) xi[a] = Bi[(ti + 16 * a) * P + cc];
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) {
            if (b2 < nb) {
                const double xj = Bj[min(tj + 16 * b2, CH_NB - 1) * P + cc];
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[a][b2] = fma(xi[a], xj, acc[a][b2]);
            }
        }
    }
    double* T = Lv + seg_toff[seg];
    const int tnr = seg_tnr[seg], tcol0 = seg_tcol0[seg];
    const int* relp = rel + seg_relptr[seg];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ti + 16 * a;  // below-row index of s
        if (i >= rb) continue;
        const int trow = relp[i - j0];
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) {
            const int j = j0 + tj + 16 * b2;
            if (j < j1 && j <= i) {
                const int tcol = rows[nc + j] - tcol0;
                T[trow + (long long)tcol * tnr] -= acc[a][b2];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// device: triangular solves, one CTA per supernode
// ------------------------------------------------------------------------------------------------
__global__ void gather_perm_kernel(int m, const int* __restrict__ perm, const double* __restrict__ src,
                                   double* __restrict__ dst, int inverse) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (!inverse) dst[i] = src[perm[i]];   // y = P b
    else dst[perm[i]] = src[i];            // x = P' y
}

// Triangular solve with a diagonal block staged in shared memory (column-major, pitch SN_SP), 128
// "unknown" threads (tid < 128 own v), four 32x32 sub-blocks: one warp per sub-block with shuffles
// (lane = row), then a rank-32 update of the other unknowns.  Same scheme as trsv_diag_kernel.
constexpr int SN_SP = 130;
constexpr int SN_SOLVE_SMEM = (CH_NB * SN_SP + 2 * CH_NB) * 8;

__device__ __forceinline__ double snode_tri_solve(const double* S, double* xs, double v, double di, int nc,
                                                  bool transposed) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!transposed) {
        for (int sb = 0; sb < CH_NB / 32; ++sb) {
            const int base = 32 * sb;
            if (base >= nc) break;
            if (warp == sb) {
#pragma unroll 8
                for (int cc = 0; cc < 32; ++cc) {
                    const double xc = __shfl_sync(0xffffffffu, v * di, cc);
                    if (lane == cc) v = xc;
                    else if (lane > cc && base + lane < nc) v = fma(-S[(base + lane) + (base + cc) * SN_SP], xc, v);
                }
                xs[base + lane] = v;
            }
            __syncthreads();
            if (tid >= base + 32 && tid < nc) {
                double acc = 0.0;
#pragma unroll 8
                for (int cc = 0; cc < 32; ++cc) acc = fma(S[tid + (base + cc) * SN_SP], xs[base + cc], acc);
                v -= acc;
            }
        }
    } else {
        for (int sb = CH_NB / 32 - 1; sb >= 0; --sb) {
            const int base = 32 * sb;
            if (base >= nc) continue;
            if (warp == sb) {
#pragma unroll 8
                for (int cc = 31; cc >= 0; --cc) {
                    const double xc = __shfl_sync(0xffffffffu, v * di, cc);
                    if (lane == cc) v = xc;
                    else if (lane < cc && base + cc < nc)
                        v = fma(-S[(base + cc) + (base + lane) * SN_SP], xc, v);
                }
                xs[base + lane] = v;
            }
            __syncthreads();
            if (tid < base) {
                double acc = 0.0;
#pragma unroll 8
                for (int cc = 0; cc < 32; ++cc)
                    if (base + cc < nc) acc = fma(S[(base + cc) + tid * SN_SP], xs[base + cc], acc);
                v -= acc;
            }
        }
    }
    return v;
}

// forward: x_s <- L_ss^-1 x_s ; x[rows below] -= B x_s.   One CTA (256 threads) per supernode.
__global__ void __launch_bounds__(256)
snode_fwd_kernel(const double* __restrict__ Lv, long long off, int nr, int nc, int col0,
                 const int* __restrict__ rows, const double* __restrict__ dinv, double* __restrict__ x) {
    extern __shared__ double sm[];
    double* S = sm;
    double* xs = S + CH_NB * SN_SP;
    const int tid = threadIdx.x;
    const double* blk = Lv + off;
#pragma unroll 4
    for (int idx = tid; idx < nc * nc; idx += 256) {
        const int cc = idx / nc, r = idx - cc * nc;
        if (r > cc) S[r + cc * SN_SP] = blk[r + (long long)cc * nr];
    }
    double v = 0.0, di = 1.0;
    if (tid < nc) {
        v = x[col0 + tid];
        di = dinv[col0 + tid];
    }
    __syncthreads();
    v = snode_tri_solve(S, xs, v, di, nc, false);  // all 256 threads walk the same barriers
    if (tid < nc) {
        x[col0 + tid] = v;
        xs[CH_NB + tid] = v;
    }
    __syncthreads();
    const double* xf = xs + CH_NB;
    for (int i = nc + tid; i < nr; i += 256) {
        double acc = 0.0;
#pragma unroll 8
        for (int cc = 0; cc < nc; ++cc) acc = fma(blk[i + (long long)cc * nr], xf[cc], acc);
        x[rows[i]] -= acc;
    }
}

// backward: x_s <- L_ss^-T (x_s - B' x[rows below]).
__global__ void __launch_bounds__(256)
snode_bwd_kernel(const double* __restrict__ Lv, long long off, int nr, int nc, int col0,
                 const int* __restrict__ rows, const double* __restrict__ dinv, double* __restrict__ x) {
    extern __shared__ double sm[];
    double* S = sm;
    double* xs = S + CH_NB * SN_SP;
    double* dots = xs + CH_NB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* blk = Lv + off;
#pragma unroll 4
    for (int idx = tid; idx < nc * nc; idx += 256) {
        const int cc = idx / nc, r = idx - cc * nc;
        if (r > cc) S[r + cc * SN_SP] = blk[r + (long long)cc * nr];
    }
    for (int cc = warp; cc < nc; cc += 8) {
        double acc = 0.0;
#pragma unroll 4
        for (int i = nc + lane; i < nr; i += 32) acc = fma(blk[i + (long long)cc * nr], x[rows[i]], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) dots[cc] = acc;
    }
    __syncthreads();
    double v = 0.0, di = 1.0;
    if (tid < nc) {
        v = x[col0 + tid] - dots[tid];
        di = dinv[col0 + tid];
    }
    v = snode_tri_solve(S, xs, v, di, nc, true);
    if (tid < nc) x[col0 + tid] = v;
}

// ---- dataflow supernodal solve: one cooperative launch per sweep -------------------------------
// CTAs take supernodes round-robin in elimination order.  Forward (owner computes): supernode t waits,
// source by source, for the descendants s that have rows inside t's columns, subtracts
// B_s[rows in t, :] y_s from its right-hand side in shared memory (fixed source order => bitwise
// reproducible, no atomics), multiplies by W_t = L_tt^-1 and publishes y_t with an epoch-numbered flag.
// Backward: supernode s waits for the supernodes that own its below-diagonal rows, forms
// y_s - B_s' z[rows], multiplies by W_s' and publishes.  Replaces 2 * nsuper single-CTA launches.
constexpr int SD_THREADS = 256;
constexpr int SD_SMEM = (CH_NB * 129 + 4 * CH_NB) * 8;

// W_s = L_ss^-1 for every supernode, one CTA per supernode, thread j = column j (forward substitution).
__global__ void __launch_bounds__(CH_NB)
snode_trtri_kernel(const double* __restrict__ Lv, const long long* __restrict__ off,
                   const int* __restrict__ first, const int* __restrict__ nrs,
                   const double* __restrict__ dinv_g, const long long* __restrict__ woff,
                   double* __restrict__ W) {
    constexpr int P = 129;
    extern __shared__ double S[];   // L strictly below the diagonal at S[r + c*P]; W' on/above it
    double* dv = S + CH_NB * P;
    const int s = blockIdx.x, j = threadIdx.x;
    const int col0 = first[s], nc = first[s + 1] - col0, nr = nrs[s];
    const double* blk = Lv + off[s];
    for (int idx = j; idx < nc * nc; idx += CH_NB) {
        const int cc = idx / nc, r = idx - cc * nc;
        if (r > cc) S[r + cc * P] = blk[r + (long long)cc * nr];
    }
    dv[j] = (j < nc) ? dinv_g[col0 + j] : 1.0;
    __syncthreads();
    if (j < nc) {
        double* w = S + j;
        w[j * P] = dv[j];
        for (int r = j + 1; r < nc; ++r) {
            double s0 = 0.0, s1 = 0.0;
            int cc = j;
            for (; cc + 1 < r; cc += 2) {
                s0 = fma(S[r + cc * P], w[cc * P], s0);
                s1 = fma(S[r + (cc + 1) * P], w[(cc + 1) * P], s1);
            }
            if (cc < r) s0 = fma(S[r + cc * P], w[cc * P], s0);
            w[r * P] = -(s0 + s1) * dv[r];
        }
    }
    __syncthreads();
    double* Wg = W + woff[s];   // column-major nc x nc, zero above the diagonal
    for (int idx = j; idx < nc * nc; idx += CH_NB) {
        const int cc = idx / nc, r = idx - cc * nc;
        Wg[idx] = (r >= cc) ? S[cc + r * P] : 0.0;
    }
}

__device__ __forceinline__ void sd_wait(volatile int* flag, int epoch) {
    if (threadIdx.x == 0) {
        while (*flag != epoch) {
        }
        __threadfence();
    }
    __syncthreads();
}

struct SolveDesc {
    const double* Lv;
    const double* W;
    const long long* off;
    const long long* woff;
    const int* first;
    const int* nr;
    const int* rowptr;
    const int* rows;
    const int* inptr;
    const int* in_s;
    const int* in_j0;
    const int* in_j1;
    const int* segptr;
    const int* seg_tid;
    int nsuper;
};

constexpr int SD_WP = 129;  // pitch of the staged W block

__global__ void __launch_bounds__(SD_THREADS)
snode_solve_dataflow_kernel(SolveDesc d, double* __restrict__ x, int* __restrict__ flags, int epoch,
                            int transposed) {
    extern __shared__ double sm[];
    double* Ws = sm;                          // W_s staged as Ws[r + c*SD_WP]
    double* rhs = Ws + CH_NB * SD_WP;         // right-hand side of this supernode's columns (nc)
    double* ys = rhs + CH_NB;                 // the published piece of another supernode
    double* part = ys + CH_NB;                // 2 x 128 partial sums
    const int tid = threadIdx.x, t = tid & 127, half = tid >> 7;
    for (int q = blockIdx.x; q < d.nsuper; q += gridDim.x) {
        const int s = transposed ? d.nsuper - 1 - q : q;
        const int col0 = d.first[s], nc = d.first[s + 1] - col0, nr = d.nr[s];
        const double* Wb = d.W + d.woff[s];
        __syncthreads();
        // stage W_s and the right-hand side while the dependencies are still in flight
        for (int idx = tid; idx < nc * nc; idx += SD_THREADS) {
            const int cc = idx / nc, r = idx - cc * nc;
            Ws[r + cc * SD_WP] = Wb[idx];
        }
        if (tid < CH_NB) rhs[tid] = (tid < nc) ? x[col0 + tid] : 0.0;
        __syncthreads();
        if (!transposed) {
            for (int e = d.inptr[s]; e < d.inptr[s + 1]; ++e) {
                const int src = d.in_s[e], j0 = d.in_j0[e], j1 = d.in_j1[e];
                const int scol0 = d.first[src], snc = d.first[src + 1] - scol0, snr = d.nr[src];
                const double* B = d.Lv + d.off[src] + snc;      // B(j, c) = B[j + c*snr]
                const int* srows = d.rows + d.rowptr[src] + snc;
                const int nj = j1 - j0;                         // <= 128 rows of src land in my columns
                const int c_lo = half * ((snc + 1) / 2), c_hi = half ? snc : (snc + 1) / 2;
                // fetch my slice of B before spinning on the source's flag
                double breg[64];
#pragma unroll
                for (int cc = 0; cc < 64; ++cc)
                    breg[cc] = (t < nj && c_lo + cc < c_hi) ? B[(j0 + t) + (long long)(c_lo + cc) * snr] : 0.0;
                sd_wait(flags + src, epoch);
                if (tid < CH_NB) ys[tid] = (tid < snc) ? __ldcg(x + scol0 + tid) : 0.0;
                __syncthreads();
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int cc = 0; cc < 64; cc += 2) {
                    a0 = fma(breg[cc], ys[min(c_lo + cc, CH_NB - 1)], a0);
                    a1 = fma(breg[cc + 1], ys[min(c_lo + cc + 1, CH_NB - 1)], a1);
                }
                part[half * CH_NB + t] = a0 + a1;
                __syncthreads();
                if (tid < nj) rhs[srows[j0 + tid] - col0] -= part[tid] + part[CH_NB + tid];
                __syncthreads();
            }
            // y_s = W rhs  (lower triangular matvec from shared memory, two halves of the columns)
            {
                const int c_lo = half * 64, c_hi = min(nc, c_lo + 64);
                double a0 = 0.0, a1 = 0.0;
                if (t < nc) {
                    int cc = c_lo;
                    for (; cc + 1 < c_hi; cc += 2) {
                        a0 = fma(Ws[t + cc * SD_WP], rhs[cc], a0);
                        a1 = fma(Ws[t + (cc + 1) * SD_WP], rhs[cc + 1], a1);
                    }
                    if (cc < c_hi) a0 = fma(Ws[t + cc * SD_WP], rhs[cc], a0);
                }
                part[half * CH_NB + t] = a0 + a1;
            }
            __syncthreads();
            if (tid < nc) x[col0 + tid] = part[tid] + part[CH_NB + tid];
        } else {
            // y_s - B_s' z, one outgoing segment (= one ancestor) at a time, farthest ancestor first:
            // thread t owns column t; its slice of the segment's rows is fetched before the wait
            const double* blk = d.Lv + d.off[s] + nc;           // B(j, c) = blk[j + c*nr]
            const int* R = d.rows + d.rowptr[s] + nc;
            for (int e = d.segptr[s + 1] - 1; e >= d.segptr[s]; --e) {
                const int tgt = d.seg_tid[e];
                // segment bounds were stored per source in seg_j0/seg_j1; recover them from the rows:
                // rows of this segment are those owned by tgt
                const int tcol0 = d.first[tgt], tcol1 = d.first[tgt + 1];
                // binary search is avoided: in_j0/in_j1 of (tgt <- s) are the same numbers, but indexed by
                // target; the outgoing copy lives in d.in_* only per target, so scan (segments are short)
                int j0 = 0, j1 = 0;
                {
                    const int rb = nr - nc;
                    int lo = 0;
                    while (lo < rb && R[lo] < tcol0) ++lo;
                    int hi = lo;
                    while (hi < rb && R[hi] < tcol1) ++hi;
                    j0 = lo;
                    j1 = hi;
                }
                const int nj = j1 - j0;
                const int r_lo = half * ((nj + 1) / 2), r_hi = half ? nj : (nj + 1) / 2;
                double breg[64];
#pragma unroll
                for (int rr = 0; rr < 64; ++rr)
                    breg[rr] = (t < nc && r_lo + rr < r_hi) ? blk[(j0 + r_lo + rr) + (long long)t * nr] : 0.0;
                sd_wait(flags + tgt, epoch);
                if (tid < CH_NB) ys[tid] = (tid < nj) ? __ldcg(x + R[j0 + tid]) : 0.0;
                __syncthreads();
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int rr = 0; rr < 64; rr += 2) {
                    a0 = fma(breg[rr], ys[min(r_lo + rr, CH_NB - 1)], a0);
                    a1 = fma(breg[rr + 1], ys[min(r_lo + rr + 1, CH_NB - 1)], a1);
                }
                part[half * CH_NB + t] = a0 + a1;
                __syncthreads();
                if (tid < nc) rhs[tid] -= part[tid] + part[CH_NB + tid];
                __syncthreads();
            }
            // z_s = W' rhs
            {
                const int c_lo = half * 64, c_hi = min(nc, c_lo + 64);
                double a0 = 0.0, a1 = 0.0;
                if (t < nc) {
                    int cc = c_lo;
                    for (; cc + 1 < c_hi; cc += 2) {
                        a0 = fma(Ws[cc + t * SD_WP], rhs[cc], a0);
                        a1 = fma(Ws[(cc + 1) + t * SD_WP], rhs[cc + 1], a1);
                    }
                    if (cc < c_hi) a0 = fma(Ws[cc + t * SD_WP], rhs[cc], a0);
                }
                part[half * CH_NB + t] = a0 + a1;
            }
            __syncthreads();
            if (tid < nc) x[col0 + tid] = part[tid] + part[CH_NB + tid];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile int*>(flags + s) = epoch;
    }
}

static int sparse_configure(nes_ctx* c) {
    static bool done = false;
    if (done) return 0;
    NES_CUDA(c, cudaFuncSetAttribute(snode_potrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SN_DIAG_SMEM));
    NES_CUDA(c, cudaFuncSetAttribute(snode_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SN_TR_SMEM));
    NES_CUDA(c, cudaFuncSetAttribute(snode_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SN_UP_SMEM));
    NES_CUDA(c, cudaFuncSetAttribute(snode_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (CH_NB * 129 + CH_NB) * 8));
    NES_CUDA(c, cudaFuncSetAttribute(snode_solve_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     SD_SMEM));
    NES_CUDA(c, cudaFuncSetAttribute(snode_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SN_SOLVE_SMEM));
    NES_CUDA(c, cudaFuncSetAttribute(snode_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SN_SOLVE_SMEM));
    done = true;
    return 0;
}

int sparse_factorize(nes_ctx* c, nes_matrix* A, nes_factor* L) {
    SparseFactor* sf = L->sparse;
    if (!sf) return fail(c, NES_ERR_INVALID, "sparse factor was not analyzed");
    NES_TRY(sparse_configure(c));
    const MatrixBase* b = A->base;
    L->factorized = 0;
    {
        StageTimer t(c, NES_STAGE_FORM);
        NES_CUDA(c, cudaMemsetAsync(sf->d_L, 0, (size_t)sf->lsize * sizeof(double), c->stream));
        NES_CUDA(c, cudaMemsetAsync(sf->d_info, 0, 2 * sizeof(int), c->stream));
        if (sf->anz > 0) {
            sparse_assemble_kernel<<<(unsigned)((sf->anz + 255) / 256), 256, 0, c->stream>>>(
                sf->anz, sf->d_ei, sf->d_ej, sf->d_edest, b->d_rowptr, b->d_colidx, b->d_csr_val,
                A->d_theta, sf->d_L);
            NES_CHECK_LAUNCH(c);
        }
    }
    {
        StageTimer t(c, NES_STAGE_FACTOR);
        for (int s = 0; s < sf->nsuper; ++s) {
            const int col0 = sf->first[s], nc = sf->first[s + 1] - col0, nr = sf->nr[s];
            const long long off = sf->off[s];
            snode_potrf_kernel<<<1, 256, SN_DIAG_SMEM, c->stream>>>(sf->d_L, off, nr, nc, col0, sf->d_dinv,
                                                                   c->dbound, sf->d_info);
            NES_CHECK_LAUNCH(c);
            const int rb = nr - nc;
            if (rb <= 0) continue;
            snode_trsm_kernel<<<(rb + SN_TR_ROWS - 1) / SN_TR_ROWS, 256, SN_TR_SMEM, c->stream>>>(
                sf->d_L, off, nr, nc, col0, sf->d_dinv);
            NES_CHECK_LAUNCH(c);
            const int nseg = sf->segptr[s + 1] - sf->segptr[s];
            if (nseg > 0) {
                dim3 grid((rb + SN_UP_ROWS - 1) / SN_UP_ROWS, nseg);
                snode_update_kernel<<<grid, 256, SN_UP_SMEM, c->stream>>>(
                    sf->d_L, off, nr, nc, sf->d_rows + sf->rowptr[s], sf->segptr[s], sf->d_seg_toff,
                    sf->d_seg_tnr, sf->d_seg_tcol0, sf->d_seg_j0, sf->d_seg_j1, sf->d_seg_relptr, sf->d_rel);
                NES_CHECK_LAUNCH(c);
            }
        }
    }
    int info[2] = {0, 0};
    NES_TRY(download(c, info, sf->d_info, sizeof(info)));
    if (info[0] != 0) {
        c->status = NES_NOT_POSDEF;
        c->minor = info[1];  // column in the permuted ordering, like cholmod_factor.minor
        return NES_NOT_POSDEF;
    }
    c->minor = sf->m;
    L->factorized = 1;
    {   // block inverses for the dataflow solve (all supernodes in one launch)
        StageTimer t(c, NES_STAGE_FACTOR);
        snode_trtri_kernel<<<sf->nsuper, CH_NB, (CH_NB * 129 + CH_NB) * 8, c->stream>>>(
            sf->d_L, sf->d_off, sf->d_first, sf->d_nr, sf->d_dinv, sf->d_woff, sf->d_W);
        NES_CHECK_LAUNCH(c);
    }
    return 0;
}

int sparse_solve_inplace(nes_ctx* c, nes_factor* L, double* d_x) {
    SparseFactor* sf = L->sparse;
    StageTimer t(c, NES_STAGE_SOLVE);
    const int m = sf->m;
    gather_perm_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(m, sf->d_perm, d_x, sf->d_x, 0);
    NES_CHECK_LAUNCH(c);
    if (sf->d_W && sf->d_flags) {
        SolveDesc d{sf->d_L, sf->d_W, sf->d_off, sf->d_woff, sf->d_first, sf->d_nr, sf->d_rowptr, sf->d_rows,
                    sf->d_inptr, sf->d_in_s, sf->d_in_j0, sf->d_in_j1, sf->d_segptr, sf->d_seg_tid, sf->nsuper};
        const int grid = sf->nsuper < c->num_sms ? sf->nsuper : c->num_sms;
        double* xp = sf->d_x;
        int* fl = sf->d_flags;
        for (int transposed = 0; transposed < 2; ++transposed) {
            int ep = ++sf->flag_epoch;
            void* args[] = {(void*)&d, (void*)&xp, (void*)&fl, (void*)&ep, (void*)&transposed};
            NES_CUDA(c, cudaLaunchCooperativeKernel((const void*)snode_solve_dataflow_kernel, dim3(grid),
                                                    dim3(SD_THREADS), args, SD_SMEM, c->stream));
            ++c->launches;
        }
        gather_perm_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(m, sf->d_perm, sf->d_x, d_x, 1);
        NES_CHECK_LAUNCH(c);
        return 0;
    }
    for (int s = 0; s < sf->nsuper; ++s) {
        const int col0 = sf->first[s], nc = sf->first[s + 1] - col0;
        snode_fwd_kernel<<<1, 256, SN_SOLVE_SMEM, c->stream>>>(sf->d_L, sf->off[s], sf->nr[s], nc, col0,
                                                  sf->d_rows + sf->rowptr[s], sf->d_dinv, sf->d_x);
        NES_CHECK_LAUNCH(c);
    }
    for (int s = sf->nsuper - 1; s >= 0; --s) {
        const int col0 = sf->first[s], nc = sf->first[s + 1] - col0;
        snode_bwd_kernel<<<1, 256, SN_SOLVE_SMEM, c->stream>>>(sf->d_L, sf->off[s], sf->nr[s], nc, col0,
                                                  sf->d_rows + sf->rowptr[s], sf->d_dinv, sf->d_x);
        NES_CHECK_LAUNCH(c);
    }
    gather_perm_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(m, sf->d_perm, sf->d_x, d_x, 1);
    NES_CHECK_LAUNCH(c);
    return 0;
}

// expand the supernodal factor to a dense lower-triangular matrix (testing) + permutation
int sparse_factor_to_dense(nes_ctx* c, nes_factor* L, double* Lout, size_t ld, int* perm_out) {
    SparseFactor* sf = L->sparse;
    std::vector<double> h((size_t)sf->lsize);
    std::vector<int> rows(sf->rowptr[sf->nsuper]), perm(sf->m);
    NES_TRY(download(c, h.data(), sf->d_L, h.size() * sizeof(double)));
    NES_TRY(download(c, rows.data(), sf->d_rows, rows.size() * sizeof(int)));
    NES_TRY(download(c, perm.data(), sf->d_perm, perm.size() * sizeof(int)));
    for (size_t j = 0; j < (size_t)sf->m; ++j)
        for (size_t i = 0; i < (size_t)sf->m; ++i) Lout[i + j * ld] = 0.0;
    for (int s = 0; s < sf->nsuper; ++s) {
        const int col0 = sf->first[s], nc = sf->first[s + 1] - col0, nr = sf->nr[s];
        const int* R = rows.data() + sf->rowptr[s];
        for (int cc = 0; cc < nc; ++cc)
            for (int r = cc; r < nr; ++r)
                Lout[(size_t)R[r] + (size_t)(col0 + cc) * ld] = h[(size_t)sf->off[s] + r + (size_t)cc * nr];
    }
    if (perm_out)
        for (int i = 0; i < sf->m; ++i) perm_out[i] = perm[i];
    return 0;
}

}  // namespace nes
