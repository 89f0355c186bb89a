// K1 + K2 for a dense constraint matrix: normal-matrix formation and blocked right-looking Cholesky.
//
// Replaces, for dense A, what the reference obtains from CHOLMOD in solve-dense
// (sparse-cholesky.lisp:409-431): cholmod_factorize on an unsymmetric B = A diag(s) forms B B' and
// factorizes it as one dense supernode (LAPACK dpotrf).  Failure protocol is the reference's:
// a non-positive pivot sets status = 1 (CHOLMOD_NOT_POSDEF) and records the column ("minor");
// the Lisp maps any non-zero status to NIL (:420-421).
//
// Per 128-column inner panel:
//   potrf_diag_kernel   one CTA: the 128x128 diagonal block arrives by one TMA box; 8-column sub-panels, every
//                       row-owning thread factors the 8x8 pivot block redundantly in registers and solves its
//                       row, rank-8 update on a 16x16 thread grid (potrf_block.cuh); TMA store
//   trsm_panel_kernel   X L_kk' = B for a 64-row slab: operands by TMA, 4 threads per row (true substitution,
//                       no explicit inverse), TMA store
//   dmma_nt_kernel      updates C -= P P' on the FP64 tensor cores (lower tiles only)
// Panels are grouped into outer panels (512 / 256 / 128 columns by remaining size) with look-ahead: see
// dense_cholesky() below; the multi-GPU variant is dense_cholesky_dist_steps().
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dmma_nt.cuh"
#include "dmma_nt64.cuh"
#include "nes_internal.h"
#include "potrf_block.cuh"
#include "trtri_block.cuh"

namespace nes {

constexpr int CH_DIAG_SMEM = (CH_NB * CH_P + CH_NB) * 8 + 128 + 16;

__global__ void __launch_bounds__(256)
potrf_diag_kernel(const __grid_constant__ CUtensorMap mapBlk, int j0, int jb,
                  double* __restrict__ dinv_out, double dbound, int* __restrict__ info, int brows) {
    // batched mode: blockIdx.y = problem, its rows start at blockIdx.y * brows, own {status, minor}
    const int brow = blockIdx.y * brows;
    info += 2 * blockIdx.y;
    dinv_out += brow;
    // dynamic shared memory of a kernel without static __shared__ starts at the CTA window base,
    // which satisfies the 128B alignment TMA needs; keeping S a plain shared pointer lets ptxas
    // use 32-bit shared addressing in the hot loops.
    extern __shared__ __align__(128) double S[];
    double* dinv = S + CH_NB * CH_P;
    uint64_t* bar = reinterpret_cast<uint64_t*>(dinv + CH_NB);
    const int tid = threadIdx.x;

    // one TMA box brings the whole 128x128 diagonal block (rows/cols past m are zero-filled)
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, CH_NB * CH_NB * 8);
        tma_load_2d(S, &mapBlk, j0 + brow, j0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    potrf_block_smem(S, dinv, jb, dbound, info, j0);

    if (tid < jb) dinv_out[j0 + tid] = dinv[tid];
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        tma_store_2d(&mapBlk, j0 + brow, j0, S);  // clipped at the matrix edge by the tensor map
        tma_store_commit_and_wait();
    }
}

// X L' = B for a 64-row slab of the panel below a full 128x128 diagonal block.
constexpr int TR_ROWS = 64;
constexpr int TR_SMEM = (CH_NB * CH_NB + TR_ROWS * CH_NB + CH_NB) * 8 + 16;

constexpr int TR_THREADS = 256;  // 4 threads per row (trsm_slab_smem)

// Operands arrive by TMA: the 128x128 diagonal block (dense box = the Ls layout, upper triangle never
// read) and the 64x128 slab of the panel (box lands as Xs[p*64 + row], rows past the matrix edge are
// zero-filled); the solved slab leaves by one TMA store (clipped at the edge).  Thread-issued staging of
// these 192 KB cost ~12 us per CTA, more than the substitution itself.
// Row mapping: slab s covers 64 rows of block s*64 / bh of the set {row_start + b*stride .. + bh}, b = 0, 1, ...
// (a contiguous panel is one block with bh >= its height; the row pieces of a P x Q distribution are the
// block rows one rank owns, `stride` = P * nbo rows apart).  Rows >= m_end are not touched.
__global__ void __launch_bounds__(TR_THREADS)
trsm_panel_kernel(const __grid_constant__ CUtensorMap mapBlk, const __grid_constant__ CUtensorMap mapSlab,
                  int j0, int row_start, int bh, int stride, int m_end, const double* __restrict__ dinv_g,
                  int brows) {
    extern __shared__ __align__(128) double sm[];
    double* Ls = sm;                       // Ls[c + p*128] = L[c][p]
    double* Xs = Ls + CH_NB * CH_NB;       // Xs[p*64 + row]
    double* dv = Xs + TR_ROWS * CH_NB;
    uint64_t* bar = reinterpret_cast<uint64_t*>(dv + CH_NB);
    const int tid = threadIdx.x;
    const int brow = blockIdx.y * brows;   // batched mode: this problem's rows
    const int off = blockIdx.x * TR_ROWS;
    const int blk = (bh > m_end) ? 0 : off / bh;
    const int row0 = row_start + blk * stride + (off - blk * bh);
    const int bend = (bh > m_end) ? m_end : min(m_end, row_start + blk * stride + bh);
    const int nrows = min(TR_ROWS, bend - row0);
    if (nrows <= 0) return;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, (CH_NB * CH_NB + TR_ROWS * CH_NB) * 8);
        tma_load_2d(Ls, &mapBlk, brow + j0, j0, bar);
        tma_load_2d(Xs, &mapSlab, brow + row0, j0, bar);
    }
    if (tid < CH_NB) dv[tid] = dinv_g[brow + j0 + tid];
    __syncthreads();
    mbar_wait(bar, 0);
    trsm_slab_smem(Ls, Xs, dv, CH_NB, nrows);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        // rows of the slab past the matrix edge are clipped by the tensor map; in batched mode a slab
        // never crosses into the next problem (row stride is a multiple of 128) and the problem's own
        // padding rows are stored back unchanged.  Rows between nrows and 64 inside the matrix (a slab
        // that ends at m_end < m) were loaded and are stored back unchanged.
        tma_store_2d(&mapSlab, brow + row0, j0, Xs);
        tma_store_commit_and_wait();
    }
}

constexpr int TI_SMEM_FWD = (CH_NB * 129 + CH_NB) * 8;
__global__ void trtri_diag_kernel(const double* __restrict__ M, long long ld, int m,
                                  const double* __restrict__ dinv_g, double* __restrict__ Winv);

static int chol_configure(nes_ctx* c) {
    static PerDeviceOnce once;
    int dev;
    if (!once.begin(&dev)) return 0;
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CH_DIAG_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(trtri_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TI_SMEM_FWD);
    once.finish(dev, e == cudaSuccess);
    NES_CUDA(c, e);
    return 0;
}

// Outer panel width by remaining size R = m - j0 (see dense_cholesky): while the trailing matrix is large
// the updates bound the step (wide panels: fewer passes over C, K = 1024 / 512 / 256: the tile's
// read-modify-write epilogue is amortised over more tensor work), once it is small the panel chain does (128
// columns: no inner narrow updates).  Measured at m = 32768: 368.0 ms with 512-column panels throughout the
// large phase, 352.7 ms with 1024-column panels while R > 8192 (354.0 with R > 4096, 359.0 with R > 24576);
// m = 16384: 50.5 -> 49.7 ms; m = 8192 is untouched (10.1 ms if 1024-column panels were used there).
static int chol_width(int m, int j0) {
    static int t1024 = -1, t512 = -1, t256 = -1;
    if (t512 < 0) {
        t1024 = 8192;  // NES_CHOL_SCHED = "t1024,t512,t256" (or the last two) overrides the thresholds
        t512 = 6144;
        t256 = 3072;
        if (const char* e = getenv("NES_CHOL_SCHED")) {
            int a = 0, b = 0, c3 = 0;
            const int n = sscanf(e, "%d,%d,%d", &a, &b, &c3);
            if (n == 3) {
                t1024 = a;
                t512 = b;
                t256 = c3;
            } else if (n == 2) {
                t512 = a;
                t256 = b;
            }
        }
    }
    const int R = m - j0;
    const int w = R > t1024 ? 1024 : (R > t512 ? 512 : (R > t256 ? 256 : CH_NB));
    return R < w ? R : w;
}

static bool chol_lookahead(const nes_ctx* c, int m) { return c->stream_aux && m > 1024; }

// ---- deferred formation (single GPU, experimental, OFF by default) ---------------------------------
// Once the trailing matrix is small the factorization is bound by its panel chain (diagonal block + TRSM,
// ~65 us per 128 columns) and most SMs idle, while the formation before it is pure tensor-core work.  With
// NES_CHOL_DEFER=d the last d columns of M are not formed up front: the region starts at zero and block
// column J+2 ("strip") receives A diag(theta) A' on the side stream during step J, between the
// rest-updates, which also accumulate into it (sums commute; everything that writes a tile is ordered on
// one stream).  A strip has fewer tiles than the GPU has SMs, so every tile is cut into k-ranges that are
// summed in a fixed order by the last CTA to arrive (dmma_nt split_all: bitwise reproducible).  The strip
// kernels leave NES_CHOL_RESERVE SMs (default 8) to the panel chain on the high-priority stream.
// Measured on B200 at m=8192, n=16384 (tools/probe_defer.py, NES_CHOL_TRACE timelines in DESIGN.md):
// form+factor 41.4 ms without, 41.8 / 42.6 ms with d = 2048 / 4096.  The strips run at 80-93% of the
// up-front kernel's rate (k-split reduction, reserved SMs) and serialise with the rest-updates on the side
// stream, which costs more than the ~3.8 ms of idle SM time they fill.  Kept as an opt-in path with its
// parity tests; the default (0) forms everything up front.
static int defer_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int defer_plan(nes_ctx* c, nes_factor* L, size_t n) {
    if (L->defer_planned && L->defer_n == n) return 0;
    cudaStreamSynchronize(c->stream);
    dev_free(c, L->d_defer_tiles);
    dev_free(c, L->d_defer_ws);
    dev_free(c, L->d_defer_counters);
    L->d_defer_tiles = nullptr;
    L->d_defer_ws = nullptr;
    L->d_defer_counters = nullptr;
    L->defer_strips.clear();
    L->defer_split = 0;
    L->defer_ntiles_a = 0;
    L->defer_planned = 1;
    L->defer_n = n;
    const int m = (int)L->m;
    const int defer = defer_env("NES_CHOL_DEFER", 0);
    const int min_m = defer_env("NES_CHOL_DEFER_MIN_M", 5120);
    if (L->dist || defer <= 0 || m < min_m || !chol_lookahead(c, m)) return 0;
    std::vector<int> bounds;  // panel boundaries p_0 = 0 < p_1 < ... < m
    for (int j0 = 0; j0 < m; j0 += chol_width(m, j0)) bounds.push_back(j0);
    // the first deferred block column must be the "J+2" of some step: p_k with k >= 2
    int split = 0;
    for (size_t k = 2; k < bounds.size(); ++k)
        if (bounds[k] >= m - defer) {
            split = bounds[k];
            break;
        }
    if (split <= 0) return 0;
    const int tm = (m + NT_BM - 1) / NT_BM, cs = split / NT_BM;
    std::vector<int2> tiles;
    // up-front tiles (block columns < cs) in the band order of nt_tile_coords: bands of NT_BAND tile rows,
    // each walked column by column, so a wave of CTAs shares operand rows through L2
    for (int r0 = 0; r0 < tm; r0 += NT_BAND) {
        const int r1 = std::min(tm, r0 + NT_BAND);
        for (int bj = 0; bj < std::min(r1, cs); ++bj)
            for (int bi = std::max(r0, bj); bi < r1; ++bi) tiles.push_back(make_int2(bi, bj));
    }
    L->defer_ntiles_a = (int)tiles.size();
    const int kchunks = ((int)n + NT_BK - 1) / NT_BK;
    const int reserve = defer_env("NES_CHOL_RESERVE", 8);
    const int G = std::max(1, c->num_sms - std::max(0, reserve));
    size_t ws_tiles = 0;
    int max_tiles = 0;
    for (size_t k = 0; k < bounds.size(); ++k) {
        if (bounds[k] < split) continue;
        nes_factor::DeferStrip st;
        st.col0 = bounds[k];
        st.first = (int)tiles.size();
        const int c1 = std::min(m, st.col0 + chol_width(m, st.col0));
        for (int bj = st.col0 / NT_BM; bj * NT_BM < c1; ++bj)
            for (int bi = bj; bi < tm; ++bi) tiles.push_back(make_int2(bi, bj));
        st.ntiles = (int)tiles.size() - st.first;
        // k-ranges per tile: the count that fills whole waves of the G CTAs best (smallest on ties)
        int best = 1;
        double best_eff = 0.0;
        const int smax = std::max(1, std::min(16, kchunks / 8));
        for (int sp = 1; sp <= smax; ++sp) {
            const double x = (double)st.ntiles * sp / G;
            const double eff = x / std::ceil(x);
            if (eff > best_eff + 1e-9) {
                best_eff = eff;
                best = sp;
            }
        }
        st.split = best;
        if (st.split > 1) ws_tiles = std::max(ws_tiles, (size_t)st.ntiles * st.split);
        max_tiles = std::max(max_tiles, st.ntiles);
        L->defer_strips.push_back(st);
    }
    L->d_defer_tiles = static_cast<int2*>(dev_alloc(c, (tiles.size() + 1) * sizeof(int2)));
    if (!L->d_defer_tiles) return c->status;
    NES_TRY(upload(c, L->d_defer_tiles, tiles.data(), tiles.size() * sizeof(int2)));
    if (ws_tiles > 0) {
        L->d_defer_ws = static_cast<double*>(dev_alloc(c, ws_tiles * NT_BM * NT_BN * sizeof(double)));
        L->d_defer_counters = static_cast<int*>(dev_alloc(c, (max_tiles + 1) * sizeof(int)));
        if (!L->d_defer_ws || !L->d_defer_counters) return c->status;
        NES_CUDA(c, cudaMemsetAsync(L->d_defer_counters, 0, (max_tiles + 1) * sizeof(int), c->stream));
    }
    L->defer_split = split;
    return 0;
}

// strip starting at column col0: M[col0.., col0..col0+w) += A diag(theta) A' on the side stream
static int defer_form_strip(nes_ctx* c, const nes_matrix* A, nes_factor* L, int col0) {
    const nes_factor::DeferStrip* st = nullptr;
    for (const auto& s : L->defer_strips)
        if (s.col0 == col0) st = &s;
    if (!st) return fail(c, NES_ERR_INVALID, "deferred formation: no strip at column %d", col0);
    const MatrixBase* b = A->base;
    NtArgs a{};
    a.C = L->d_M;
    a.ldc = (long long)L->ld;
    a.M = a.N = (int)L->m;
    a.K = (int)b->n;
    a.scale = A->d_theta;
    a.alpha = 1.0;
    a.beta = 1.0;
    a.same_operand = 1;
    a.tile_list = L->d_defer_tiles + st->first;
    a.ntiles = st->ntiles;
    if (st->split > 1) {
        a.split_all = 1;
        a.split_s = st->split;
        a.split_ws = L->d_defer_ws;
        a.split_counters = L->d_defer_counters;
    }
    const int G = std::max(1, c->num_sms - std::max(0, defer_env("NES_CHOL_RESERVE", 8)));
    cudaError_t e = nt_launch(b->map, b->map, a, G, c->stream_aux);
    ++c->launches;
    if (e != cudaSuccess)
        return fail(c, NES_ERR_CUDA, "deferred formation launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// As = A diag(s): one column per blockIdx.x, rows strided over the block and blockIdx.y (ld is a multiple of 16)
__global__ void __launch_bounds__(256)
scale_columns_kernel(const double* __restrict__ A, const double* __restrict__ s, double* __restrict__ As,
                     size_t ld) {
    const size_t col = blockIdx.x;
    const double sj = s[col];
    const double2* src = reinterpret_cast<const double2*>(A + col * ld);
    double2* dst = reinterpret_cast<double2*>(As + col * ld);
    for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < ld / 2; i += (size_t)gridDim.y * blockDim.x) {
        double2 v = src[i];
        v.x *= sj;
        v.y *= sj;
        dst[i] = v;
    }
}

// Formation operand.  The fused kernel multiplies theta into the B fragment in registers (4 DMUL per 32 DMMA
// and lane): measured 3.4% slower than the same kernel without a scale (34.7 vs 35.9 TFLOP/s at m = 16384,
// 34.5 vs 35.7 at m = 8192; tools/probe_unscaled_formation.py) -- DMUL shares the FP64 datapath with DMMA.
// One elementwise pass As = A diag(s) costs 16 m n bytes of HBM traffic (5.3 ms at m = 32768, n = 65536
// against 66 ms saved) and a second buffer the size of A; M = As As' is then literally the reference's
// formulation (scale-sparse! followed by the factorization of B B', sparse-cholesky.lisp:461-473, :408).
// Default: pre-scaled; NES_FORM_FUSED=1, or an allocation that fails, falls back to the fused kernel.
static int form_prescale(nes_ctx* c, const nes_matrix* A, nes_factor* L, bool* use) {
    *use = false;
    const MatrixBase* b = A->base;
    if (!A->d_scale || b->n == 0 || getenv("NES_FORM_FUSED") || L->defer_split > 0) return 0;
    if (!L->d_As || L->As_ld != b->ld || L->As_n != b->n) {
        cudaStreamSynchronize(c->stream);
        dev_free(c, L->d_As);
        L->d_As = nullptr;
        const int saved_status = c->status;
        L->d_As = static_cast<double*>(dev_alloc(c, b->ld * (b->n ? b->n : 1) * sizeof(double)));
        if (!L->d_As) {  // no room for the copy: the fused kernel needs none
            c->status = saved_status;
            c->err[0] = 0;
            cudaGetLastError();
            return 0;
        }
        L->As_ld = b->ld;
        L->As_n = b->n;
        if (make_operand_map(&L->mapAs, L->d_As, (long long)b->m, (long long)b->n, (long long)b->ld) != 0)
            return fail(c, NES_ERR_CUDA, "cuTensorMapEncodeTiled failed for the scaled copy of A");
        // rows m..ld of A are zero and stay zero in the copy (the kernel scales whole columns)
    }
    const unsigned gy = (unsigned)std::max<size_t>(1, std::min<size_t>(8, (b->ld / 2 + 255) / 256));
    scale_columns_kernel<<<dim3((unsigned)b->n, gy), 256, 0, c->stream>>>(b->d_val, A->d_scale, L->d_As, b->ld);
    NES_CHECK_LAUNCH(c);
    *use = true;
    return 0;
}

int dense_form_normal(nes_ctx* c, const nes_matrix* A, nes_factor* L, bool allow_defer) {
    const MatrixBase* b = A->base;
    int split = 0;
    if (allow_defer && !L->dist) {
        NES_TRY(defer_plan(c, L, b->n));
        split = L->defer_split;
    }
    StageTimer timer(c, NES_STAGE_FORM);
    bool prescaled = false;
    NES_TRY(form_prescale(c, A, L, &prescaled));
    const CUtensorMap& opmap = prescaled ? L->mapAs : b->map;
    NtArgs a{};
    a.C = L->d_M;
    a.ldc = (long long)L->ld;
    a.M = a.N = (int)b->m;
    a.rowA0 = a.rowB0 = 0;
    a.k0 = 0;
    a.K = (int)b->n;
    a.scale = prescaled ? nullptr : A->d_theta;  // nullptr: plain X X'
    a.alpha = 1.0;
    a.beta = 0.0;
    a.lower = 1;
    a.same_operand = 1;
    if (L->dist) {  // only the tiles this rank owns (block-cyclic outer block columns)
        a.tile_list = L->d_tile_list;
        a.ntiles = L->ntiles_owned;
        a.lower = 0;
    }
    {
        const double dm = (double)b->m, dd = (double)(b->m - (size_t)split);
        c->form_flops = (split > 0 ? dm * dm - dd * dd : dm * dm) * (double)b->n;
    }
    if (split > 0) {
        // block columns >= split start at zero and are formed inside dense_cholesky
        const size_t d = b->m - (size_t)split;
        NES_CUDA(c, cudaMemset2DAsync(L->d_M + split + (size_t)split * L->ld, L->ld * sizeof(double), 0,
                                      d * sizeof(double), d, c->stream));
        a.tile_list = L->d_defer_tiles;
        a.ntiles = L->defer_ntiles_a;
        a.lower = 0;
    }
    // balance the partial last wave of the static tile round-robin (6% at m=8192 on 148 SMs)
    {
        const int tm = (a.M + NT_BM - 1) / NT_BM;
        const int ntiles = a.tile_list ? a.ntiles : tm * (tm + 1) / 2;
        const int grid = ntiles < c->num_sms ? ntiles : c->num_sms;
        int sr = 0, ss = 0;
        nt_plan_split(ntiles, (a.K + NT_BK - 1) / NT_BK, grid, &sr, &ss);
        if (sr > 0) {
            if (!c->d_split_counters) {
                c->d_split_counters = static_cast<int*>(dev_alloc(c, 1024 * sizeof(int)));
                if (!c->d_split_counters) return c->status;
                NES_CUDA(c, cudaMemsetAsync(c->d_split_counters, 0, 1024 * sizeof(int), c->stream));
            }
            a.split_ws = ensure_ws(c, WS_SPLIT, (size_t)sr * ss * NT_BM * NT_BN * sizeof(double));
            if (!a.split_ws) return c->status;
            a.split_counters = c->d_split_counters;
            a.split_r = sr;
            a.split_s = ss;
        }
    }
    cudaError_t e = nt_launch(opmap, opmap, a, c->num_sms, c->stream);
    ++c->launches;
    if (e != cudaSuccess)
        return fail(c, NES_ERR_CUDA, "formation kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// Diagonal block + TRSM of one 128-column inner panel for `nbatch` stacked problems (brows = row
// stride between problems; nbatch = 1, brows = 0 for a single matrix).  Used by batch.cu.
int chol_panel_launch(nes_ctx* c, const CUtensorMap& mapBlk, const CUtensorMap& mapSlab, int i0, int ib,
                      int m, double* dinv, int* info, int nbatch, int brows) {
    NES_TRY(chol_configure(c));
    potrf_diag_kernel<<<dim3(1, nbatch), 256, CH_DIAG_SMEM, c->stream>>>(mapBlk, i0, ib, dinv, c->dbound,
                                                                        info, brows);
    NES_CHECK_LAUNCH(c);
    const int rest = m - i0 - ib;
    if (rest > 0) {
        trsm_panel_kernel<<<dim3((rest + TR_ROWS - 1) / TR_ROWS, nbatch), TR_THREADS, TR_SMEM, c->stream>>>(
            mapBlk, mapSlab, i0, i0 + CH_NB, 1 << 30, 0, m, dinv, brows);
        NES_CHECK_LAUNCH(c);
    }
    return 0;
}

// Which kernel applies the Cholesky updates: the 128 x 128 one (default) or, with NES_UPDATE_KERNEL=64, the
// 128 x 64 half-tile kernel with two CTAs per SM (dmma_nt64.cuh).  With 512-column panels the half-tile kernel is
// 4% faster on the factorization at m = 32768 (368 -> 354 ms); 1024-column panels give the 128 x 128 kernel the
// same gain and the two do not add up, so it is not the default here.  (Its early versions lost updates -- a
// stage released before its fragment loads had returned; found and fixed, see ptx_util.cuh: mbar_arrive_after
// and DESIGN.md section 5.)
static bool update_uses_nt64() {
    const char* e = getenv("NES_UPDATE_KERNEL");
    return e && atoi(e) == 64;
}

// One dmma_nt launch: C[r0.., c0..c0+ncols) -= X[r0.., k0..k0+K) X[c0..c0+ncols, k0..k0+K)^T
static int chol_update(nes_ctx* c, nes_factor* L, int r0, int c0, int nrows, int ncols, int k0, int K,
                       int lower, cudaStream_t stream = nullptr, bool one_tile_per_cta = false) {
    NtArgs a{};
    a.C = L->d_M + r0 + (long long)c0 * (long long)L->ld;
    a.ldc = (long long)L->ld;
    a.M = nrows;
    a.N = ncols;
    a.rowA0 = r0;
    a.rowB0 = c0;
    a.k0 = k0;
    a.K = K;
    a.scale = nullptr;
    a.alpha = -1.0;
    a.beta = 1.0;
    a.lower = lower;
    a.same_operand = (r0 == c0) ? 1 : 0;
    // one_tile_per_cta: a non-persistent grid, so SMs come free every tile and a higher-priority stream
    // (the panel factorization of the look-ahead) gets them at tile granularity
    cudaError_t e = update_uses_nt64()
                        ? nt64_launch(L->mapM, L->mapM68, a, one_tile_per_cta ? 0 : 2 * c->num_sms,
                                      stream ? stream : c->stream)
                        : nt_launch(L->mapM, L->mapM, a, one_tile_per_cta ? (1 << 30) : c->num_sms,
                                    stream ? stream : c->stream);
    ++c->launches;
    if (e != cudaSuccess)
        return fail(c, NES_ERR_CUDA, "Cholesky update launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// Two-level blocking: outer panels of NBO columns, factored left-looking in 128-column inner
// panels (narrow DMMA update, diagonal block, TRSM), then one wide trailing update with K = NBO.
// A larger K halves (NBO=256) or quarters (512) the number of passes over the trailing matrix and
// the epilogue share of each tile.
// Inverses of the 128x128 diagonal blocks of L, one CTA per block, thread j = column j of W = L_ii^-1
// by forward substitution.  They serve the SOLVE phase only (dataflow TRSV): there a block solve is on
// the critical path of every block row, and a triangular matvec with W is ~10x shorter than a
// substitution.  The error it adds, cond(L_ii) eps, is far below the cond(M) eps of the solve itself;
// the factorization (TRSM) never uses inverses, so the ||LL' - M|| gate is untouched.
constexpr int TI_P = 129;
constexpr int TI_SMEM = (CH_NB * TI_P + CH_NB) * 8;

__global__ void __launch_bounds__(TRTRI_THREADS)
trtri_diag_kernel(const double* __restrict__ M, long long ld, int m, const double* __restrict__ dinv_g,
                  double* __restrict__ Winv) {
    // one shared array holds both triangles: L strictly below the diagonal at S[r + c*P] (r > c) and
    // W' on and above it, W(r, j) at S[j + r*P] (r >= j); four lanes per column (trtri_block.cuh)
    extern __shared__ double S[];
    double* dv = S + CH_NB * TI_P;
    const int tid = threadIdx.x;
    const int j0 = blockIdx.x * CH_NB;
    const int jb = min(CH_NB, m - j0);
    for (int idx = tid; idx < jb * jb; idx += TRTRI_THREADS) {
        const int cc = idx / jb, r = idx - cc * jb;
        if (r > cc) S[r + cc * TI_P] = M[(j0 + r) + (long long)(j0 + cc) * ld];
    }
    if (tid < CH_NB) dv[tid] = (tid < jb) ? dinv_g[j0 + tid] : 1.0;
    __syncthreads();
    trtri_columns_smem<TI_P>(S, dv, jb);
    __syncthreads();
    double* Wg = Winv + (size_t)blockIdx.x * CH_NB * CH_NB;  // column-major 128 x 128, zero upper part
    for (int idx = tid; idx < CH_NB * CH_NB; idx += TRTRI_THREADS) {
        const int cc = idx >> 7, r = idx & 127;
        Wg[idx] = (r >= cc && r < jb && cc < jb) ? S[cc + r * TI_P] : 0.0;
    }
}

__global__ void info_to_minor_kernel(int* info) {
    // info = {status, minor}  ->  {status, status ? minor : INT_MAX} so MIN over ranks finds the first
    if (threadIdx.x == 0 && info[0] == 0) info[1] = 0x7fffffff;
}

// ---- distributed factorization (P x Q block-cyclic, panels pipelined in row chunks) ---------------------
// nes_dist.cu plans, for every panel (block column J), the fixed sequence of messages it travels in: row
// chunks (and, for P > 1, the diagonal block first).  Streams on every rank:
//   stream    (high)  the message that carries the diagonal block, when this rank is its root: update of those
//                     rows with panel J-1, the inner 128-column panels (narrow DMMA update, potrf_diag,
//                     trsm_panel), pack
//   stream_c  (mid)   the other messages this rank is the root of: update, trsm against the finished diagonal
//                     block, narrow updates, pack -- chunk by chunk
//   stream_b  (high)  communication: one ncclBroadcast per message in plan order, unpack on the receivers
//   stream_aux (low)  the trailing update with panel J of this rank's tiles in block columns >= J+2, one tile
//                     per CTA; the tiles of column J+2 go first and signal ev_colready[J+2]
// The root of a message of panel J+1 waits only for the message(s) of panel J that carry the same rows and
// block row J+1 (absolute chunk boundaries make that one message), so the owner of the next panel starts on
// its first chunk while the later chunks of the current one are still being solved and sent, and the
// panel chain per step is diagonal block + first chunk + one small broadcast instead of whole panel +
// whole-panel broadcast.  Every rank issues the same collectives in the same order on stream_b.
__global__ void pack_rows_kernel(double* __restrict__ M, long long ld, int m, int col0, int ncols, int row_start,
                                 int nblocks, int bh, int stride, double* __restrict__ buf, int ldp,
                                 double* __restrict__ dinv, int unpack) {
    const long long body = (long long)ldp * ncols, total = body + (dinv ? ncols : 0);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        if (idx >= body) {
            const int cidx = (int)(idx - body);
            if (unpack) dinv[col0 + cidx] = buf[idx];
            else buf[idx] = dinv[col0 + cidx];
            continue;
        }
        const int cidx = (int)(idx / ldp), r = (int)(idx - (long long)cidx * ldp);
        const int blk = r / bh;
        const int row = row_start + blk * stride + (r - blk * bh);
        if (blk >= nblocks || row >= m) continue;
        double* src = M + row + (long long)(col0 + cidx) * ld;
        if (unpack) *src = buf[idx];
        else buf[idx] = *src;
    }
}

static int dist_setup(nes_ctx* c, nes_factor* L) {
    DistPlan& pl = *L->dist;
    if (pl.ev_start) return 0;
    auto mk = [&](cudaEvent_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess; };
    bool ok = mk(&pl.ev_start) && mk(&pl.ev_end[0]) && mk(&pl.ev_end[1]) && mk(&pl.ev_end[2]);
    const int nmsg = pl.msg_base[pl.nblk];
    pl.ev_arrived.assign(nmsg, nullptr);
    pl.ev_packed.assign(nmsg, nullptr);
    pl.ev_colready.assign(pl.nblk, nullptr);
    pl.ev_diagdone.assign(pl.nblk, nullptr);
    for (int i = 0; i < nmsg && ok; ++i) ok = mk(&pl.ev_arrived[i]) && mk(&pl.ev_packed[i]);
    for (int i = 0; i < pl.nblk && ok; ++i) ok = mk(&pl.ev_colready[i]) && mk(&pl.ev_diagdone[i]);
    if (!ok) return fail(c, NES_ERR_CUDA, "cudaEventCreate failed for the distributed factorization");
    for (int i = 0; i < DistPlan::kStages; ++i) {
        pl.d_stage[i] = static_cast<double*>(dev_alloc(c, (pl.max_msg_doubles + 16) * sizeof(double)));
        if (!pl.d_stage[i]) return c->status;
    }
    return 0;
}

// this rank's tiles [tile_begin, tile_end) -= panel(k0, K) panel(k0, K)'
static int dist_update(nes_ctx* c, nes_factor* L, int k0, int K, int tile_begin, int tile_end, cudaStream_t stream,
                       bool one_tile_per_cta) {
    if (tile_begin >= tile_end || K <= 0) return 0;
    NtArgs a{};
    a.C = L->d_M;
    a.ldc = (long long)L->ld;
    a.M = a.N = (int)L->m;
    a.rowA0 = a.rowB0 = 0;
    a.k0 = k0;
    a.K = K;
    a.alpha = -1.0;
    a.beta = 1.0;
    a.same_operand = 1;
    a.tile_list = L->d_tile_list + tile_begin;
    a.ntiles = tile_end - tile_begin;
    // one_tile_per_cta: SMs come free at tile granularity for the panel chain on the higher-priority streams.
    // NES_REST_TPC = tiles per CTA of these launches (default 1; measured at m = 32768 on 2 GPUs: 204 ms per
    // factorization with 1 or 2, 208 with 4; one persistent CTA per SM is 266 ms -- the chain starves -- and is
    // not offered: its residual at m = 32768 was 1e-7 even with every stream serialised, an open defect of
    // dmma_nt's beta != 0 path at > 100 tiles per CTA that the one-tile launches do not exercise).
    const char* ev = getenv("NES_REST_TPC");
    int tpc = ev ? atoi(ev) : 1;
    if (tpc < 1) tpc = 1;
    int max_ctas = c->num_sms;
    if (one_tile_per_cta) max_ctas = std::max(c->num_sms, (a.ntiles + tpc - 1) / tpc);
    // The distributed schedule keeps the 128 x 128 kernel by default: every number in DESIGN.md section 5 was
    // validated with it at 2, 4 and 8 GPUs.  NES_DIST_KERNEL_DEBUG=1/2/3 selects the half-tile kernel for the
    // trailing updates / the panel chain / both: since the stage-release fix (ptx_util.cuh: mbar_arrive_after) it
    // is exact in every replay (tools/debug_dist_single.py on one GPU, 2 real GPUs at m = 32768: 203 -> 197 ms)
    // and only waits for an 8-GPU validation run to become the default.
    bool use64 = false;
    if (const char* dbg = getenv("NES_DIST_KERNEL_DEBUG")) {
        const int v = atoi(dbg);
        const bool is_rest = (stream == c->stream_aux);
        use64 = (v == 1) ? is_rest : (v == 2 ? !is_rest : v == 3);
    }
    cudaError_t e = use64 ? nt64_launch(L->mapM, L->mapM68, a, 0, stream)
                          : nt_launch(L->mapM, L->mapM, a, max_ctas, stream);
    ++c->launches;
    if (e != cudaSuccess)
        return fail(c, NES_ERR_CUDA, "distributed update launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// the inner 128-column panels of panel `pn` on the rows of message `d` (root only)
static int dist_factor_rows(nes_ctx* c, nes_factor* L, const DistPanel& pn, const DistMsg& d, cudaStream_t stream) {
    const int m = (int)L->m;
    for (int t = 0; t * CH_NB < pn.jbo; ++t) {
        const int i0 = pn.j0 + t * CH_NB;
        const int ib = (m - i0 < CH_NB) ? m - i0 : CH_NB;
        if (t > 0) NES_TRY(dist_update(c, L, pn.j0, t * CH_NB, d.seg[t], d.seg[t + 1], stream, true));
        if (d.has_diag) {
            potrf_diag_kernel<<<1, 256, CH_DIAG_SMEM, stream>>>(L->mapBlk, i0, ib, L->d_dinv, c->dbound, L->d_info, 0);
            NES_CHECK_LAUNCH(c);
            const int end = std::min(m, d.row_start + d.bh);   // contiguous rows below the diagonal block
            const int rest = end - i0 - ib;
            if (rest > 0) {
                trsm_panel_kernel<<<(rest + TR_ROWS - 1) / TR_ROWS, TR_THREADS, TR_SMEM, stream>>>(
                    L->mapBlk, L->mapSlab, i0, i0 + CH_NB, 1 << 30, 0, end, L->d_dinv, 0);
                NES_CHECK_LAUNCH(c);
            }
        } else {
            const int slabs = d.nblocks * (d.bh / TR_ROWS);
            trsm_panel_kernel<<<slabs, TR_THREADS, TR_SMEM, stream>>>(L->mapBlk, L->mapSlab, i0, d.row_start, d.bh,
                                                                     d.stride, m, L->d_dinv, 0);
            NES_CHECK_LAUNCH(c);
        }
    }
    return 0;
}

static int dist_pack(nes_ctx* c, nes_factor* L, const DistPanel& pn, const DistMsg& d, double* buf, int unpack,
                     cudaStream_t stream) {
    const long long total = (long long)d.rows * pn.jbo + pn.jbo;
    int grid = (int)std::min<long long>((total + 255) / 256, (long long)c->num_sms * 8);
    if (grid < 1) grid = 1;
    pack_rows_kernel<<<grid, 256, 0, stream>>>(L->d_M, (long long)L->ld, (int)L->m, pn.j0, pn.jbo, d.row_start,
                                               d.nblocks, d.bh, d.stride, buf, d.rows,
                                               d.has_diag ? L->d_dinv : nullptr, unpack);
    NES_CHECK_LAUNCH(c);
    return 0;
}

// NES_CHOL_TRACE=1: CUDA-event timeline of the single-GPU factorization (which kernel ran when, on which
// stream), printed to stderr after a synchronisation.  Debugging aid; off by default.
struct CholTrace {
    struct Span {
        const char* what;
        int col;
        cudaEvent_t a, b;
    };
    bool on = false;
    int rank = 0;
    cudaEvent_t base = nullptr;
    std::vector<Span> spans;
    void begin(cudaStream_t s) {
        on = getenv("NES_CHOL_TRACE") != nullptr;
        if (!on) return;
        cudaEventCreate(&base);
        cudaEventRecord(base, s);
    }
    int open(const char* what, int col, cudaStream_t s) {
        if (!on) return -1;
        Span sp{what, col, nullptr, nullptr};
        cudaEventCreate(&sp.a);
        cudaEventCreate(&sp.b);
        cudaEventRecord(sp.a, s);
        spans.push_back(sp);
        return (int)spans.size() - 1;
    }
    void close(int id, cudaStream_t s) {
        if (id >= 0) cudaEventRecord(spans[id].b, s);
    }
    void dump(cudaStream_t s) {
        if (!on) return;
        cudaStreamSynchronize(s);
        for (auto& sp : spans) {
            float t0 = 0.f, t1 = 0.f;
            cudaEventElapsedTime(&t0, base, sp.a);
            cudaEventElapsedTime(&t1, base, sp.b);
            fprintf(stderr, "chol-trace r%d %-6s col %5d  %9.3f -> %9.3f  (%7.3f ms)\n", rank, sp.what, sp.col, t0,
                    t1, t1 - t0);
            cudaEventDestroy(sp.a);
            cudaEventDestroy(sp.b);
        }
        cudaEventDestroy(base);
        spans.clear();
    }
};

static int dense_cholesky_dist_steps(nes_ctx* c, nes_factor* L, CholTrace& tr) {
    NES_TRY(dist_setup(c, L));
    DistPlan& pl = *L->dist;
    const int nblk = pl.nblk, me = c->rank, P = pl.P, tpb = pl.tpb;
    // NES_DIST_PAIR=1: trailing updates of two consecutive panels applied together (K = 2 nbo), see below
    const bool pair = getenv("NES_DIST_PAIR") && atoi(getenv("NES_DIST_PAIR")) != 0;
    cudaStream_t S0 = c->stream, S1 = c->stream_aux, S2 = c->stream_b, S3 = c->stream_c;
    if (const char* e = getenv("NES_DIST_SERIAL")) {  // debugging: merge streams into the main one (bit 0: trailing
        const int v = atoi(e);                        // updates, bit 1: communication, bit 2: bulk pieces; 1 = all)
        const int mask = (v == 1) ? 7 : v >> 1;
        if (mask & 1) S1 = S0;
        if (mask & 2) S2 = S0;
        if (mask & 4) S3 = S0;
    }
    NES_CUDA(c, cudaEventRecord(pl.ev_start, S0));
    NES_CUDA(c, cudaStreamWaitEvent(S1, pl.ev_start, 0));
    NES_CUDA(c, cudaStreamWaitEvent(S2, pl.ev_start, 0));
    NES_CUDA(c, cudaStreamWaitEvent(S3, pl.ev_start, 0));
    for (int J = 0; J < nblk; ++J) {
        const DistPanel& pn = pl.panels[J];
        for (size_t k = 0; k < pn.msgs.size(); ++k) {
            const DistMsg& d = pn.msgs[k];
            const int gi = pl.msg_base[J] + (int)k;
            double* stage = pl.d_stage[gi % DistPlan::kStages];
            const size_t count = (size_t)d.rows * pn.jbo + (d.has_diag ? pn.jbo : 0);
            if (d.root == me) {
                cudaStream_t st = d.urgent ? S0 : S3;
                int t = -1;
                if (J > 0) {
                    const DistPanel& pv = pl.panels[J - 1];
                    if (d.dep >= 0) NES_CUDA(c, cudaStreamWaitEvent(st, pl.ev_arrived[pl.msg_base[J - 1] + d.dep], 0));
                    if (J >= 2) NES_CUDA(c, cudaStreamWaitEvent(st, pl.ev_colready[J], 0));
                    t = tr.open(d.urgent ? "upd0" : "upd", pn.j0, st);
                    NES_TRY(dist_update(c, L, pv.j0, pv.jbo, d.seg[0], d.seg[tpb], st, true));
                    tr.close(t, st);
                }
                if (!d.has_diag)
                    NES_CUDA(c, cudaStreamWaitEvent(st, P > 1 ? pl.ev_arrived[pl.msg_base[J] + pn.diag_msg]
                                                              : pl.ev_diagdone[J], 0));
                t = tr.open(d.has_diag ? "diag" : "rows", pn.j0, st);
                NES_TRY(dist_factor_rows(c, L, pn, d, st));
                tr.close(t, st);
                if (d.has_diag) NES_CUDA(c, cudaEventRecord(pl.ev_diagdone[J], st));
                if (gi >= DistPlan::kStages)  // the staging slot's previous broadcast is over
                    NES_CUDA(c, cudaStreamWaitEvent(st, pl.ev_arrived[gi - DistPlan::kStages], 0));
                if (!(c->nranks == 1 && getenv("NES_DIST_NOPACK")))  // debugging (one GPU: nobody reads the stage)
                    NES_TRY(dist_pack(c, L, pn, d, stage, 0, st));
                NES_CUDA(c, cudaEventRecord(pl.ev_packed[gi], st));
                NES_CUDA(c, cudaStreamWaitEvent(S2, pl.ev_packed[gi], 0));
                t = tr.open("send", pn.j0, S2);
                if (c->nranks > 1) NES_TRY(dist_broadcast(c, stage, count, d.root, S2));
                tr.close(t, S2);
            } else {
                int t = tr.open("recv", pn.j0, S2);
                NES_TRY(dist_broadcast(c, stage, count, d.root, S2));
                NES_TRY(dist_pack(c, L, pn, d, stage, 1, S2));
                tr.close(t, S2);
            }
            NES_CUDA(c, cudaEventRecord(pl.ev_arrived[gi], S2));
        }
        // trailing update with panel J: block columns >= J+2 (column J+1 is updated by the roots of its messages)
        if (J + 2 < nblk) {
            NES_CUDA(c, cudaStreamWaitEvent(S1, pl.ev_arrived[pl.msg_base[J + 1] - 1], 0));
            const DistPanel& p2 = pl.panels[J + 2];
            int t = tr.open("rest", pn.j0, S1);
            NES_TRY(dist_update(c, L, pn.j0, pn.jbo, p2.col_begin, p2.col_end, S1, true));
            NES_CUDA(c, cudaEventRecord(pl.ev_colready[J + 2], S1));
            if (!pair) {
                NES_TRY(dist_update(c, L, pn.j0, pn.jbo, p2.col_end, L->ntiles_owned, S1, true));
            } else if ((J & 1) == 0) {
                // paired updates: an even panel only reaches the next TWO block columns on its own ...
                if (J + 3 < nblk) {
                    const DistPanel& p3 = pl.panels[J + 3];
                    NES_TRY(dist_update(c, L, pn.j0, pn.jbo, p3.col_begin, p3.col_end, S1, true));
                }
            } else {
                // ... and travels on with its odd successor: one update with K = two panels (contiguous columns)
                // for everything from block column J+3 on -- half the passes over the trailing matrix, and
                // the tile's read-modify-write epilogue amortised over twice the tensor work
                const DistPanel& pv = pl.panels[J - 1];
                NES_TRY(dist_update(c, L, pv.j0, pv.jbo + pn.jbo, p2.col_end, L->ntiles_owned, S1, true));
            }
            tr.close(t, S1);
        }
    }
    NES_CUDA(c, cudaEventRecord(pl.ev_end[0], S1));
    NES_CUDA(c, cudaEventRecord(pl.ev_end[1], S2));
    NES_CUDA(c, cudaEventRecord(pl.ev_end[2], S3));
    for (int i = 0; i < 3; ++i) NES_CUDA(c, cudaStreamWaitEvent(S0, pl.ev_end[i], 0));
    return 0;
}

int dense_cholesky(nes_ctx* c, nes_factor* L, const nes_matrix* A) {
    StageTimer timer(c, NES_STAGE_FACTOR);
    NES_TRY(chol_configure(c));
    const int m = (int)L->m;
    const long long ld = (long long)L->ld;
    const int P = L->dist ? 2 : 1;  // 2 = the distributed schedule (also under NES_FORCE_DIST on one rank)
    NES_CUDA(c, cudaMemsetAsync(L->d_info, 0, 2 * sizeof(int), c->stream));
    if (P > 1) {
        CholTrace tr;
        tr.rank = c->rank;
        tr.begin(c->stream);
        NES_TRY(dense_cholesky_dist_steps(c, L, tr));
        tr.dump(c->stream);
    }
    if (P == 1) {
        // Look-ahead (single GPU).  After panel J: (1) `stream` brings block column J+1 up to date and
        // factors panel J+1 (latency-bound: one CTA for the diagonal block, a few for the TRSM), while
        // (2) the low-priority side stream applies panel J to the rest of the trailing matrix with one
        // tile per CTA.  Without it the ~64 diagonal blocks and TRSMs (4 ms at m = 8192) sit between the
        // updates with 147 SMs idle.  Hazards: (1) of step J+1 touches columns the rest-update J writes,
        // so `stream` waits for ev_update; rest-update J+1 needs panel J+1 (ev_panel) and follows
        // rest-update J in stream order.
        // Panel width by remaining size: chol_width().  Deferred formation (defer_plan): block column J+2
        // is formed on the side stream at the head of step J, before the rest-update that needs panel J.
        auto width = [&](int j0) { return chol_width(m, j0); };
        const int defer = (A && L->defer_planned) ? L->defer_split : 0;
        auto factor_panel = [&](int j0, int jbo) -> int {
            for (int i0 = j0; i0 < j0 + jbo; i0 += CH_NB) {
                const int ib = (m - i0 < CH_NB) ? m - i0 : CH_NB;
                if (i0 > j0)  // bring block column i0 up to date with the inner panels already factored
                    NES_TRY(chol_update(c, L, i0, i0, m - i0, ib, j0, i0 - j0, 0));
                potrf_diag_kernel<<<1, 256, CH_DIAG_SMEM, c->stream>>>(L->mapBlk, i0, ib, L->d_dinv,
                                                                      c->dbound, L->d_info, 0);
                NES_CHECK_LAUNCH(c);
                const int rest = m - i0 - ib;
                if (rest > 0) {
                    trsm_panel_kernel<<<(rest + TR_ROWS - 1) / TR_ROWS, TR_THREADS, TR_SMEM, c->stream>>>(
                        L->mapBlk, L->mapSlab, i0, i0 + CH_NB, 1 << 30, 0, m, L->d_dinv, 0);
                    NES_CHECK_LAUNCH(c);
                }
            }
            return 0;
        };
        const bool lookahead = chol_lookahead(c, m);
        CholTrace tr;
        tr.begin(c->stream);
        if (defer > 0) {  // the side stream starts after the up-front formation and the zero fill
            NES_CUDA(c, cudaEventRecord(c->ev_panel, c->stream));
            NES_CUDA(c, cudaStreamWaitEvent(c->stream_aux, c->ev_panel, 0));
        }
        int jbo = width(0);
        NES_TRY(factor_panel(0, jbo));
        bool pending_update = false;
        for (int j0 = 0; j0 < m;) {
            const int r0 = j0 + jbo;
            if (r0 >= m) break;
            const int jbn = width(r0);
            if (!lookahead) {
                NES_TRY(chol_update(c, L, r0, r0, m - r0, m - r0, j0, jbo, 1));
                NES_TRY(factor_panel(r0, jbn));
                j0 = r0;
                jbo = jbn;
                continue;
            }
            const int r1 = r0 + jbn;
            NES_CUDA(c, cudaEventRecord(c->ev_panel, c->stream));                  // panel J is final
            if (pending_update) NES_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_update, 0));
            int t = tr.open("colupd", r0, c->stream);
            NES_TRY(chol_update(c, L, r0, r0, m - r0, jbn, j0, jbo, 0));           // (1) block column J+1
            tr.close(t, c->stream);
            if (r1 < m) {                                                          // (2) the rest, side stream
                if (defer > 0 && r1 >= defer) {
                    t = tr.open("strip", r1, c->stream_aux);
                    NES_TRY(defer_form_strip(c, A, L, r1));
                    tr.close(t, c->stream_aux);
                }
                NES_CUDA(c, cudaStreamWaitEvent(c->stream_aux, c->ev_panel, 0));
                t = tr.open("rest", r1, c->stream_aux);
                NES_TRY(chol_update(c, L, r1, r1, m - r1, m - r1, j0, jbo, 1, c->stream_aux, true));
                tr.close(t, c->stream_aux);
                NES_CUDA(c, cudaEventRecord(c->ev_update, c->stream_aux));
                pending_update = true;
            }
            t = tr.open("panel", r0, c->stream);
            NES_TRY(factor_panel(r0, jbn));
            tr.close(t, c->stream);
            j0 = r0;
            jbo = jbn;
        }
        if (pending_update) NES_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_update, 0));
        tr.dump(c->stream);
    }
    if (L->d_Winv) {  // block inverses for the solve phase (off the factorization's critical path)
        trtri_diag_kernel<<<(m + CH_NB - 1) / CH_NB, TRTRI_THREADS, TI_SMEM, c->stream>>>(L->d_M, ld, m, L->d_dinv,
                                                                                  L->d_Winv);
        NES_CHECK_LAUNCH(c);
    }
    if (c->nranks > 1) {  // a failed pivot is only seen by the owner of its panel: agree on {status, first minor}
        info_to_minor_kernel<<<1, 32, 0, c->stream>>>(L->d_info);
        NES_CHECK_LAUNCH(c);
        NES_TRY(dist_allreduce_int(c, L->d_info, 1, 1));
        NES_TRY(dist_allreduce_int(c, L->d_info + 1, 1, 0));
    }
    int info[2] = {0, 0};
    NES_TRY(download(c, info, L->d_info, sizeof(info)));
    if (info[0] != 0) {
        c->status = NES_NOT_POSDEF;
        c->minor = info[1];
        L->factorized = 0;
        return NES_NOT_POSDEF;
    }
    c->minor = m;
    L->factorized = 1;
    return 0;
}

}  // namespace nes
