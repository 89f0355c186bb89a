// K3/K4/K5 for a sparse (CSC) constraint matrix: supernodal MULTIFRONTAL Cholesky of A diag(theta) A'.
//
// Replaces what sparse-newton-solve.lisp / affine-scaling.lisp get from CHOLMOD:
//   cholmod_analyze    (sparse-cholesky.lisp:509, affine-scaling.lisp:270-271)  -> sparse_analyze  (host: sparse_symbolic.cu)
//   cholmod_factorize  (sparse-cholesky.lisp:512, 543)                          -> sparse_factorize (GPU)
//   cholmod_solve/2    (sparse-cholesky.lisp:515, 546)                          -> sparse_solve_inplace
// and the counters printed by the reference (anz, aatfl, lnz, fl; affine-scaling.lisp:273-279).
//
// Numeric phase.  The assembly tree is processed level by level (nested dissection makes the levels
// wide: hundreds of independent supernodes at the bottom); every level is three launches over all of
// its supernodes:
//   mf_potrf_kernel   one CTA per supernode: diagonal block to shared memory, extend-add of the children's
//                     update matrices into it, in-smem Cholesky (potrf_block.cuh), write back
//   mf_trsm_kernel    one CTA per 64-row slab of the rows below: slab to shared memory, extend-add,
//                     X L' = B by substitution, write back
//   mf_syrk_kernel    persistent, two CTAs per SM, 128 x 64 tiles of the update matrix
//                     U_s = -L21 L21' on the FP64 tensor cores (DMMA m8n8k4, operands by TMA from
//                     per-supernode tensor maps, 2-stage mbarrier pipeline), then the extend-add of the
//                     children's update matrices into the tile (one CTA's epilogue overlaps the other's
//                     tensor work)
// Extend-add is OWNER-COMPUTES: the CTA that owns a piece of the parent's front pulls the matching
// sub-rectangle of every child's update matrix (children in ascending order, coalesced along the
// child's columns, rows scattered through the parent-relative map), so there are no atomics and the
// factor is bitwise reproducible.  Original entries of M are assembled straight into the supernode
// blocks (one thread per entry of tril(M): merge-join of two CSR rows of A with theta).
// Update matrices live in a pool whose slots are reused as soon as the parent has consumed them.
//
// Solves are multifrontal too and level-synchronous: forward, supernode s subtracts its children's update
// VECTORS from its right-hand side, multiplies by W_s = L_ss^-1 (inverted once per factorization) and
// leaves u_s = B_s y_s + (children's entries below its columns); backward, it gathers the solution at
// its rows below, forms y_s - B_s' z and multiplies by W_s'.
//
// Multi-GPU (nranks > 1): the symbolic phase maps disjoint subtrees of the assembly tree to ranks
// (heaviest-first, flops from the symbolic counts) and leaves the top of the tree to everybody.  A rank
// factors its own subtrees (phase A), publishes the update matrices of its subtree roots with ONE
// ncclBroadcast (they are contiguous in the pool), and every rank then factors the top redundantly
// (phase B), so no further exchange is needed; the solves exchange the subtree roots' update vectors
// the same way and finish with an all-reduce of the solution pieces.
// Storage: supernode s is a dense nr x nc column-major block (ld = nr rounded up to 16) of L.
#include <algorithm>
#include <cstdlib>
#include <numeric>

#include "dmma_nt.cuh"
#include "nes_internal.h"
#include "potrf_block.cuh"
#include "trtri_block.cuh"
#include "sparse_symbolic.h"

namespace nes {

int dist_allreduce_sum(nes_ctx* c, double* d_buf, size_t count);  // nes_dist.cu

struct MfDesc {
    double* Lv;
    double* U;
    double* dinv;
    double* W;
    double* uvec;
    const long long* off;
    const long long* uoff;
    const long long* woff;
    const long long* vptr;
    const int* first;
    const int* nr;
    const int* ld;
    const int* ldu;
    const int* rowptr;
    const int* rows;
    const int* childptr;
    const int* child;
    const int* relptr;
    const int* rel;
    const int* cut;
    const int* nb0;
    const int* tbptr;
    const int* tb;
    const CUtensorMap* maps;    // per supernode block: 132 x 32 boxes
    const CUtensorMap* maps68;  // 68 x 32 boxes (B operand of the 128 x 64 SYRK tiles)
    int* info;
    double dbound;
};

// NES_SPARSE_SYNC=1: synchronise after every launch of the sparse path and name the kernel that failed
static bool sparse_sync_debug() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NES_SPARSE_SYNC");
        v = (e && *e && *e != '0') ? 1 : 0;
    }
    return v == 1;
}
#define MF_LAUNCHED(c, name)                                                                         \
    do {                                                                                             \
        NES_CHECK_LAUNCH(c);                                                                         \
        if (sparse_sync_debug()) {                                                                   \
            cudaError_t e2__ = cudaStreamSynchronize((c)->stream);                                   \
            if (e2__ != cudaSuccess)                                                                 \
                return nes::fail((c), NES_ERR_CUDA, "%s failed: %s", name, cudaGetErrorString(e2__)); \
        }                                                                                            \
    } while (0)

// launch schedule of one phase (A: the supernodes this rank owns, B: the replicated top)
struct Phase {
    std::vector<int> pptr, tptr, yptr;  // nlevels + 1 offsets into the task arrays
    std::vector<int> psplit, tsplit;    // per level: tasks [ptr, split) have nc <= 64, [split, next) are wider
    int* d_potrf = nullptr;             // supernode ids, level by level
    int2* d_trsm = nullptr;             // (supernode, 64-row slab)
    int4* d_syrk = nullptr;             // (supernode, tile row, column block of syrk_bn() columns, nc)
    std::vector<int> fptr, bptr;        // nlevels + 1 offsets into the solve task arrays
    int2* d_ftail = nullptr;            // forward sweep: (supernode, 512-row chunk of its rows below)
    int2* d_bdots = nullptr;            // backward sweep: (supernode, group of 16 columns)
    int count = 0;
};

struct SparseFactor {
    Symbolic S;
    int m = 0;
    MfDesc d{};
    Phase phase[2];
    // device copies of the symbolic arrays
    std::vector<void*> owned;  // everything to free
    int* d_perm = nullptr;
    int* d_ei = nullptr;
    int* d_ej = nullptr;
    long long* d_edest = nullptr;
    int* d_owner = nullptr;
    int* d_all = nullptr;      // supernodes factored on this rank (phase A then phase B)
    int nall = 0;
    double* d_x = nullptr;
    double* d_dots = nullptr;  // backward sweep: B_s' z per column
    int* d_info = nullptr;
    long long wsize = 0;
    // The launch sequences of a factorization (~180 launches) and of a solve (~140) depend only on the
    // pattern: they are captured into CUDA graphs on first use and replayed while their operands stay put
    // (single rank only: the multi-rank sequences contain NCCL calls).
    struct GraphCache {
        cudaGraphExec_t exec = nullptr;
        const void* k0 = nullptr;
        const void* k1 = nullptr;
        double k2 = 0.0;
        long long launches = 0;
    };
    GraphCache g_factor[4];  // keyed by the scale vector (Newton step and repair step use different ones)
    int g_factor_next = 0;
    GraphCache g_solve[4];  // keyed by the right-hand-side vector (the IPM drivers solve into 2-3 different ones)
    int g_solve_next = 0;
    bool graphs_off = false;
};

template <class F>
static int run_captured(nes_ctx* c, SparseFactor* sf, SparseFactor::GraphCache& g, const void* k0, const void* k1,
                        double k2, F enqueue) {
    static const bool env_off = getenv("NES_NO_GRAPH") != nullptr;
    if (env_off || sf->graphs_off || c->nranks > 1 || sparse_sync_debug() || c->timing > 1) return enqueue();
    if (g.exec && (g.k0 != k0 || g.k1 != k1 || g.k2 != k2)) {
        cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
    }
    if (!g.exec) {
        const long long l0 = c->launches;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            sf->graphs_off = true;
            return enqueue();
        }
        const int rc = enqueue();
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        if (rc == 0 && e == cudaSuccess && graph) e = cudaGraphInstantiate(&g.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != 0 || e != cudaSuccess || !g.exec) {  // could not capture: run directly from now on
            cudaGetLastError();
            g.exec = nullptr;
            sf->graphs_off = true;
            c->launches = l0;
            return rc != 0 ? rc : enqueue();
        }
        g.launches = c->launches - l0;
        c->launches = l0;
        g.k0 = k0;
        g.k1 = k1;
        g.k2 = k2;
    }
    NES_CUDA(c, cudaGraphLaunch(g.exec, c->stream));
    c->launches += g.launches;
    return 0;
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


// Diagonal block of every supernode of a level: load, extend-add, Cholesky, store.  `ncmax` (64 or 128,
// bounds nc over the launch) sizes the shared block, so narrow supernodes run several CTAs per SM.
// The block moves by 1D bulk copies, one per column (columns start 128-byte aligned; nc rounded up
// to even rows: the extra row is a zero pad row of the layout), all in flight at once.
static inline int mf_diag_smem(int ncmax) { return (ncmax * CH_P + CH_NB) * 8 + 16; }

__global__ void __launch_bounds__(256)
mf_potrf_kernel(const MfDesc d, const int* __restrict__ list, int ncmax) {
    extern __shared__ __align__(128) double S[];
    double* dinv = S + ncmax * CH_P;
    uint64_t* bar = reinterpret_cast<uint64_t*>(dinv + CH_NB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = list[blockIdx.x];
    const int col0 = d.first[s], nc = d.first[s + 1] - col0, ld = d.ld[s];
    const int nc2 = (nc + 1) & ~1;
    double* blk = d.Lv + d.off[s];
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, (uint32_t)(nc * nc2 * 8));
    }
    __syncthreads();
    if (warp == 0)
        for (int cc = lane; cc < nc; cc += 32) bulk_load_1d(S + cc * CH_P, blk + (long long)cc * ld, nc2 * 8, bar);
    mbar_wait(bar, 0);
    int* srl = reinterpret_cast<int*>(dinv);  // the child's parent-relative rows (dinv is written later)
    for (int q = d.childptr[s]; q < d.childptr[s + 1]; ++q) {
        const int c = d.child[q];
        const int cutc = d.cut[c];
        const double* Uc = d.U + d.uoff[c];
        const int lduc = d.ldu[c];
        if (tid < cutc) srl[tid] = d.rel[d.relptr[c] + tid];
        __syncthreads();
        // lane owns rows lane + 32k of the child's leading cutc x cutc triangle; four columns per step so
        // that sixteen independent loads are in flight per lane (the gathers are a chain of L2 round trips)
        for (int j = warp; j < cutc; j += 32) {
            double v[4][4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int jj = j + 8 * t;
                const double* u = Uc + (long long)min(jj, cutc - 1) * lduc;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = lane + 32 * k;
                    v[t][k] = (jj < cutc && i >= jj && i < cutc) ? u[i] : 0.0;
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int jj = j + 8 * t;
                if (jj < cutc) {
                    const int tcol = srl[jj] * CH_P;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = lane + 32 * k;
                        if (i >= jj && i < cutc) S[srl[i] + tcol] += v[t][k];
                    }
                }
            }
        }
        __syncthreads();
    }
    potrf_block_smem(S, dinv, nc, d.dbound, d.info, col0);
    if (tid < nc) d.dinv[col0 + tid] = dinv[tid];
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
        for (int cc = lane; cc < nc; cc += 32) bulk_store_1d(blk + (long long)cc * ld, S + cc * CH_P, nc2 * 8);
        tma_store_commit_and_wait();
    }
}

// Rows below the diagonal block, 64 at a time: load, extend-add, X L' = B, store.  Shared memory is
// sized by `ncmax` (64 or 128).  L's columns and the slab's columns arrive by bulk copies; the slab
// leaves the same way (16-row granularity: the layout pads every block to 16 rows).
constexpr int MF_TR_ROWS = 64;
constexpr int MF_LP = 132;  // pitch of L's columns in shared memory  } = 4 (mod 16): the DMMA fragment loads
constexpr int MF_XP = 68;   // pitch of the slab's columns            } of trsm_slab_smem_t are conflict-free
static inline int mf_tr_smem(int ncmax) { return (ncmax * MF_LP + ncmax * MF_XP + CH_NB) * 8 + 16; }

__global__ void __launch_bounds__(256)
mf_trsm_kernel(const MfDesc d, const int2* __restrict__ tasks, int ncmax) {
    extern __shared__ __align__(128) double sm[];
    double* Ls = sm;                        // Ls[c + p*MF_LP] = L[c][p]
    double* Xs = Ls + ncmax * MF_LP;        // Xs[p*MF_XP + row]
    double* dv = Xs + ncmax * MF_XP;
    uint64_t* bar = reinterpret_cast<uint64_t*>(dv + CH_NB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = tasks[blockIdx.x].x, slab = tasks[blockIdx.x].y;
    const int col0 = d.first[s], nc = d.first[s + 1] - col0, nr = d.nr[s], ld = d.ld[s], nb0 = d.nb0[s];
    const int nu = nr - nc;
    const int i0 = slab * MF_TR_ROWS;                 // first below-row of the slab
    const int nrows = min(MF_TR_ROWS, nu - i0);
    const int nrows16 = (nrows + 15) & ~15;           // rows moved (pad rows of the layout included)
    double* blk = d.Lv + d.off[s];
    double* slab_g = blk + nb0 + i0;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, (uint32_t)(nc * (nb0 + nrows16) * 8));
    }
    __syncthreads();
    if (warp == 0) {
        for (int p = lane; p < nc; p += 32) {
            bulk_load_1d(Ls + p * MF_LP, blk + (long long)p * ld, nb0 * 8, bar);   // rows nc..nb0-1 are zero pad
            bulk_load_1d(Xs + p * MF_XP, slab_g + (long long)p * ld, nrows16 * 8, bar);
        }
    }
    // columns nc..nb0-1 of Ls (read by the 32-column blocking of trsm_slab_smem) are zero
    for (int idx = tid; idx < (nb0 - nc) * CH_NB; idx += 256) Ls[(nc + (idx >> 7)) * MF_LP + (idx & 127)] = 0.0;
    if (tid < CH_NB) dv[tid] = (tid < nc) ? d.dinv[col0 + tid] : 1.0;
    mbar_wait(bar, 0);
    __syncthreads();
    __shared__ int srow[MF_TR_ROWS], scol[CH_NB];
    for (int q = d.childptr[s]; q < d.childptr[s + 1]; ++q) {
        const int c = d.child[q];
        const int* tb = d.tb + d.tbptr[c];
        const int ilo = tb[slab], ihi = tb[slab + 1];
        if (ihi <= ilo) continue;  // uniform over the CTA
        const int cutc = d.cut[c];
        const int* rl = d.rel + d.relptr[c];
        const double* Uc = d.U + d.uoff[c];
        const int lduc = d.ldu[c];
        __syncthreads();
        if (tid < ihi - ilo) srow[tid] = rl[ilo + tid] - (nc + i0);   // slab-local row (ihi - ilo <= 64)
        if (tid < cutc) scol[tid] = rl[tid] * MF_XP;                  // target column offset in Xs
        __syncthreads();
        const int ni = ihi - ilo;
        // lane owns child rows ilo + lane, ilo + lane + 32; eight child columns (sixteen loads) per step
        for (int j = warp; j < cutc; j += 64) {
            double v[8][2];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int jj = min(j + 8 * k, cutc - 1);
                const double* ucol = Uc + (long long)jj * lduc + ilo;
                v[k][0] = (lane < ni) ? ucol[lane] : 0.0;
                v[k][1] = (lane + 32 < ni) ? ucol[lane + 32] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int jj = j + 8 * k;
                if (jj < cutc) {
                    double* xcol = Xs + scol[jj];
                    if (lane < ni) xcol[srow[lane]] += v[k][0];
                    if (lane + 32 < ni) xcol[srow[lane + 32]] += v[k][1];
                }
            }
        }
    }
    __syncthreads();
    trsm_slab_smem_t<MF_LP, MF_XP, true>(Ls, Xs, dv, nc, nrows);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
        for (int p = lane; p < nc; p += 32) bulk_store_1d(slab_g + (long long)p * ld, Xs + p * MF_XP, nrows16 * 8);
        tma_store_commit_and_wait();
    }
}

// Update matrices of a level on the FP64 tensor cores.  Task = (supernode, tile row bi, column block cb,
// nc): the 128 x BN tile U(128 bi.., BN cb..) = -L21[rows] L21[cols]' with K = nc <= 128, plus the
// children's contributions; only tiles that touch the lower triangle are listed.
// Same pipeline as dmma_nt_kernel (thread 0 doubles as the TMA producer, DMMA warps of 64 x 32, row
// pitches 132 / 68 so the fragment loads are conflict-free); the tensor maps of the supernode's block
// come from device arrays, rows past the block and columns past nc are zero-filled by TMA.
// Extend-add happens IN REGISTERS: for every child the tile's rows and columns get an inverse map in
// shared memory (tile row -> row of the child's update matrix, from the slab table of the symbolic
// phase, no searching), and every lane gathers the child's entries that fall on its own accumulator
// elements (independent loads, no read-modify-write of global memory, one store per element).
// With K <= 128 a tile is 4 k-chunks of tensor work (18 us) followed by an epilogue of gathers and
// stores that takes twice as long, during which the DMMA pipe of a one-CTA-per-SM kernel idles.  The
// default is therefore BN = 64: 128 x 64 tiles, 4 warps, 2 stages (100 KB) -- TWO CTAs per SM, so one
// CTA's epilogue overlaps the other's mainloop.  BN = 128 (8 warps, 3 stages, one CTA per SM) is kept
// for comparison (NES_SYRK_TILE=128).
template <int BN>
struct SyrkCfg {
    static constexpr int WN = BN / 32;                 // warps along the columns
    static constexpr int WARPS = 2 * WN;
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int BPITCH = BN + 4;              // 132 or 68: == 4 mod 16
    static constexpr int B_BYTES = BPITCH * NT_BK * 8;
    static constexpr int STAGE_BYTES = NT_TILE_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 128) ? 3 : 2;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 16 + (NT_BM + BN) * 4 + 128;
    static constexpr int CTAS_PER_SM = (BN == 128) ? 1 : 2;
};

template <int BN>
__global__ void __launch_bounds__(SyrkCfg<BN>::THREADS, SyrkCfg<BN>::CTAS_PER_SM)
mf_syrk_kernel(const MfDesc d, const int4* __restrict__ tasks, int ntasks) {
    using Cfg = SyrkCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    int* inv_r = reinterpret_cast<int*>(empty + STAGES + 2);
    int* inv_c = inv_r + NT_BM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_producer = (threadIdx.x == 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const uint32_t never = (ntasks < 0) ? 0xffffffffu : 0u;  // always 0, unknown to the compiler (mbar_arrive_after)
    int p_task = blockIdx.x, p_kc = 0, p_kch = 0, p_nb0 = 0;
    int4 pt = make_int4(0, 0, 0, 0);
    uint32_t p_it = 0;
    bool p_valid = false;
    if (is_producer && p_task < ntasks) {
        pt = tasks[p_task];
        p_kch = (pt.w + NT_BK - 1) / NT_BK;
        p_nb0 = (pt.w + 31) & ~31;
        p_valid = true;
    }
    auto produce = [&]() {
        if (!p_valid) return;
        const int st_i = p_it % STAGES;
        const uint32_t ph = (p_it / STAGES) & 1;
        mbar_wait(&empty[st_i], ph ^ 1);
        const int col0 = pt.z * BN;
        const bool diag = (col0 / NT_BM == pt.y);  // the column block lies inside the row tile: one load
        uint8_t* st = smem + st_i * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[st_i], diag ? NT_TILE_BYTES : NT_TILE_BYTES + Cfg::B_BYTES);
        const CUtensorMap* mp = d.maps + pt.x;
        const CUtensorMap* mpb = (BN == 128 ? d.maps : d.maps68) + pt.x;
        if (p_kc == 0) {
            fence_tensormap_acquire(mp);
            if (BN != 128) fence_tensormap_acquire(mpb);
        }
        // (the box must start on a 16-byte boundary: an odd row coordinate of an 8-byte type is an illegal
        // instruction, measured with tools/tma_gmem_probe.cu -- the layout keeps nb0 a multiple of 32)
        tma_load_2d(st, mp, p_nb0 + pt.y * NT_BM, p_kc * NT_BK, &full[st_i]);
        if (!diag) tma_load_2d(st + NT_TILE_BYTES, mpb, p_nb0 + col0, p_kc * NT_BK, &full[st_i]);
        ++p_it;
        if (++p_kc >= p_kch) {
            p_task += gridDim.x;
            p_kc = 0;
            p_valid = p_task < ntasks;
            if (p_valid) {
                pt = tasks[p_task];
                p_kch = (pt.w + NT_BK - 1) / NT_BK;
                p_nb0 = (pt.w + 31) & ~31;
            }
        }
    };
    if (is_producer)
        for (int i = 0; i < STAGES - 1; ++i) produce();

    const int g = lane >> 2, t4 = lane & 3;
    const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
    const int a_off = t4 * NT_PITCH + wm * 64 + g;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
        const int4 tk = tasks[t];
        const int s = tk.x, bi = tk.y, col0 = tk.z * BN, nc = tk.w;
        const int kch = (nc + NT_BK - 1) / NT_BK;
        const bool diag = (col0 / NT_BM == bi);
        // B fragments: own tile (pitch BPITCH), or rows col0 - 128 bi .. of the A tile on diagonal tiles
        const int bpitch = diag ? NT_PITCH : Cfg::BPITCH;
        const int b_off = t4 * bpitch + (diag ? col0 - bi * NT_BM : 0) + wn * 32 + g;
        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kc = 0; kc < kch; ++kc, ++it) {
            if (is_producer) produce();
            __syncwarp();
            const int st_i = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(&full[st_i], ph);
            const double* sA = reinterpret_cast<const double*>(smem + st_i * Cfg::STAGE_BYTES);
            const double* sB = diag ? sA : sA + NT_PITCH * NT_BK;
            const double* ap = sA + a_off;
            const double* bp = sB + b_off;
            uint32_t dep = 0;
#pragma unroll
            for (int ks = 0; ks < NT_BK / 4; ++ks) {
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = ap[ks * 4 * NT_PITCH + i * 8];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = bp[ks * 4 * bpitch + j * 8];
                dep = frag_dependency(dep, af, bf);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            // the stage is released only once every fragment load of it has returned (ptx_util.cuh)
            if (lane == 0) mbar_arrive_after(&empty[st_i], dep, never);
        }
        const int nu = d.nr[s] - nc, ldu = d.ldu[s];
        const int nslab = (nu + 63) >> 6;
        const int lr0 = wm * 64 + g;        // my rows inside the tile: lr0 + 8 i
        const int lc0 = wn * 32 + 2 * t4;   // my columns: lc0 + 8 j + c2
        // ---- extend-add in registers: acc holds +L21 L21', children are SUBTRACTED, the store negates
        for (int q = d.childptr[s]; q < d.childptr[s + 1]; ++q) {
            const int c = d.child[q];
            const int* tb = d.tb + d.tbptr[c];
            const int jlo = tb[col0 >> 6], jhi = tb[min((col0 >> 6) + BN / 64, nslab)];
            const int ilo = tb[2 * bi], ihi = tb[min(2 * bi + 2, nslab)];
            if (ihi <= ilo || jhi <= jlo) continue;  // uniform over the CTA
            const int* rl = d.rel + d.relptr[c];
            __syncthreads();  // the previous child's maps are no longer read
            if (threadIdx.x < NT_BM) inv_r[threadIdx.x] = -1;
            if (threadIdx.x < BN) inv_c[threadIdx.x] = -1;
            __syncthreads();
            for (int i = ilo + (int)threadIdx.x; i < ihi; i += Cfg::THREADS) inv_r[rl[i] - nc - bi * NT_BM] = i;
            for (int j = jlo + (int)threadIdx.x; j < jhi; j += Cfg::THREADS) inv_c[rl[j] - nc - col0] = j;
            __syncthreads();
            const double* Uc = d.U + d.uoff[c];
            const int lduc = d.ldu[c];
            int ci[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ci[i] = inv_r[lr0 + 8 * i];
            // two batches of 32 independent loads (4 columns x 8 rows), then the subtractions: two memory
            // round trips per child instead of one per column (entries outside the child contribute 0.0,
            // which leaves the accumulator bit-identical)
#pragma unroll
            for (int jh = 0; jh < 4; jh += 2) {
                double v[2][2][8];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int cj = inv_c[lc0 + 8 * (jh + jj) + c2];
                        const double* ucol = Uc + (long long)max(cj, 0) * lduc;
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[jj][c2][i] = (cj >= 0 && ci[i] >= cj) ? ucol[ci[i]] : 0.0;
                    }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i][jh + jj][c2] -= v[jj][c2][i];
            }
        }
        // ---- store U tile = -acc (whole tile inside the matrix; only i >= j is ever read)
        double* Us = d.U + d.uoff[s];
        const int row_base = bi * NT_BM + lr0;
        const int col_base = col0 + lc0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                const int col = col_base + j * 8 + c2;
                if (col < nu) {
                    double* cp = Us + (long long)col * ldu;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row_base + i * 8;
                        if (row < nu) cp[row] = -acc[i][j][c2];
                    }
                }
            }
    }
}

// column block width of mf_syrk_kernel: 64 (two CTAs per SM) unless NES_SYRK_TILE=128
static int syrk_bn() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("NES_SYRK_TILE");
        v = (e && atoi(e) == 128) ? 128 : 64;
    }
    return v;
}

// W_s = L_ss^-1 for the listed supernodes, one CTA each (forward substitution per column, four lanes per column).
constexpr int MF_TI_SMEM = (CH_NB * 129 + CH_NB) * 8;
__global__ void __launch_bounds__(TRTRI_THREADS)
mf_trtri_kernel(const MfDesc d, const int* __restrict__ list) {
    constexpr int P = 129;
    extern __shared__ double S[];   // L strictly below the diagonal at S[r + c*P]; W' on/above it
    double* dv = S + CH_NB * P;
    const int s = list[blockIdx.x], tid = threadIdx.x;
    const int col0 = d.first[s], nc = d.first[s + 1] - col0, ld = d.ld[s];
    const double* blk = d.Lv + d.off[s];
    for (int idx = tid; idx < nc * nc; idx += TRTRI_THREADS) {
        const int cc = idx / nc, r = idx - cc * nc;
        if (r > cc) S[r + cc * P] = blk[r + (long long)cc * ld];
    }
    if (tid < CH_NB) dv[tid] = (tid < nc) ? d.dinv[col0 + tid] : 1.0;
    __syncthreads();
    trtri_columns_smem<P>(S, dv, nc);   // four lanes per column (trtri_block.cuh)
    __syncthreads();
    double* Wg = d.W + d.woff[s];   // column-major nc x nc, zero above the diagonal
    for (int idx = tid; idx < nc * nc; idx += TRTRI_THREADS) {
        const int cc = idx / nc, r = idx - cc * nc;
        Wg[idx] = (r >= cc) ? S[cc + r * P] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// device: multifrontal triangular solves, two launches per level and sweep
// ------------------------------------------------------------------------------------------------
// W_s and B_s are read straight from global memory (every element is used once).  The long part of a
// supernode -- the product with its nu x nc block B_s -- is cut into row chunks (forward) or column
// groups (backward) that run on different SMs; a single CTA per supernode was latency-bound at ~10 GB/s.
constexpr int MS_THREADS = 256;
constexpr int MS_CHUNK = 2048;
constexpr int MS_FWD_ROWS = 256;   // rows of B_s per CTA in the forward sweep (multiple of 64)
constexpr int MS_BWD_COLS = 16;    // columns of B_s per CTA in the backward sweep

// forward, head: rhs_s = b_s - (children's update vectors inside my columns); y_s = W_s rhs_s
__global__ void __launch_bounds__(MS_THREADS)
mf_fwd_head_kernel(const MfDesc d, const int* __restrict__ list, double* __restrict__ x) {
    __shared__ double rhs[CH_NB], part[2 * CH_NB];
    const int tid = threadIdx.x, t = tid & 127, half = tid >> 7;
    const int s = list[blockIdx.x];
    const int col0 = d.first[s], nc = d.first[s + 1] - col0;
    if (tid < CH_NB) rhs[tid] = (tid < nc) ? x[col0 + tid] : 0.0;
    __syncthreads();
    for (int q = d.childptr[s]; q < d.childptr[s + 1]; ++q) {
        const int c = d.child[q];
        const int cutc = d.cut[c];
        const double* uc = d.uvec + d.vptr[c];
        const int* rl = d.rel + d.relptr[c];
        for (int i = tid; i < cutc; i += MS_THREADS) rhs[rl[i]] -= uc[i];
        __syncthreads();
    }
    // y = W rhs, W lower triangular, column-major nc x nc: thread t owns row t, two halves of the columns
    const double* Wb = d.W + d.woff[s];
    const int c_lo = half * 64, c_hi = min(t + 1, min(nc, c_lo + 64));
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (t < nc) {
        int cc = c_lo;
        // sixteen loads in flight per step (W comes from L2: the product is a chain of round trips otherwise);
        // the four accumulators receive their columns in the same order as in the 4-wide loop below
        for (; cc + 15 < c_hi; cc += 16) {
            double w[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) w[u] = Wb[t + (cc + u) * nc];
#pragma unroll
            for (int u = 0; u < 16; u += 4) {
                a0 = fma(w[u + 0], rhs[cc + u + 0], a0);
                a1 = fma(w[u + 1], rhs[cc + u + 1], a1);
                a2 = fma(w[u + 2], rhs[cc + u + 2], a2);
                a3 = fma(w[u + 3], rhs[cc + u + 3], a3);
            }
        }
        for (; cc + 3 < c_hi; cc += 4) {
            a0 = fma(Wb[t + (cc + 0) * nc], rhs[cc + 0], a0);
            a1 = fma(Wb[t + (cc + 1) * nc], rhs[cc + 1], a1);
            a2 = fma(Wb[t + (cc + 2) * nc], rhs[cc + 2], a2);
            a3 = fma(Wb[t + (cc + 3) * nc], rhs[cc + 3], a3);
        }
        for (; cc < c_hi; ++cc) a0 = fma(Wb[t + cc * nc], rhs[cc], a0);
    }
    part[half * CH_NB + t] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (tid < nc) x[col0 + tid] = part[tid] + part[CH_NB + tid];
}

// forward, tail: rows [256 k, 256 k + 256) of u_s = B_s y_s + (children's entries that land there).
// Lane l of every warp owns rows l, l+32, ... (coalesced along the columns of B_s), warp w owns the
// columns w, w+8, ...; the eight partial vectors are summed through shared memory in warp order.
__global__ void __launch_bounds__(MS_THREADS)
mf_fwd_tail_kernel(const MfDesc d, const int2* __restrict__ tasks, const double* __restrict__ x) {
    __shared__ double ys[CH_NB];
    __shared__ double part[8][MS_FWD_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = tasks[blockIdx.x].x, chunk = tasks[blockIdx.x].y;
    const int col0 = d.first[s], nc = d.first[s + 1] - col0, nr = d.nr[s], ld = d.ld[s];
    const int nu = nr - nc;
    const int i0 = chunk * MS_FWD_ROWS, i1 = min(nu, i0 + MS_FWD_ROWS);
    if (tid < CH_NB) ys[tid] = (tid < nc) ? x[col0 + tid] : 0.0;
    __syncthreads();
    const double* B = d.Lv + d.off[s] + d.nb0[s] + i0;
    double a[MS_FWD_ROWS / 32];
#pragma unroll
    for (int k = 0; k < MS_FWD_ROWS / 32; ++k) a[k] = 0.0;
    // two columns (sixteen loads) per step; every a[k] still receives its columns in ascending order
    for (int cc = warp; cc < nc; cc += 16) {
        const int c2 = cc + 8;
        const double* bc = B + (long long)cc * ld;
        const double* bd = B + (long long)min(c2, nc - 1) * ld;
        const double y = ys[cc], y2 = (c2 < nc) ? ys[c2] : 0.0;
        double v[MS_FWD_ROWS / 32], v2[MS_FWD_ROWS / 32];
#pragma unroll
        for (int k = 0; k < MS_FWD_ROWS / 32; ++k) {
            const int i = lane + 32 * k;
            v[k] = (i0 + i < i1) ? bc[i] : 0.0;
            v2[k] = (i0 + i < i1 && c2 < nc) ? bd[i] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < MS_FWD_ROWS / 32; ++k) {
            const int i = lane + 32 * k;
            if (i0 + i < i1) {
                a[k] = fma(v[k], y, a[k]);
                if (c2 < nc) a[k] = fma(v2[k], y2, a[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < MS_FWD_ROWS / 32; ++k) part[warp][lane + 32 * k] = a[k];
    __syncthreads();
    double* us = d.uvec + d.vptr[s];
    for (int i = tid; i < i1 - i0; i += MS_THREADS) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += part[w][i];
        us[i0 + i] = v;
    }
    const int nslab = (nu + 63) >> 6;
    for (int q = d.childptr[s]; q < d.childptr[s + 1]; ++q) {
        __syncthreads();
        const int c = d.child[q];
        const int* tb = d.tb + d.tbptr[c];
        const int ilo = tb[min(nslab, i0 >> 6)], ihi = tb[min(nslab, (i0 + MS_FWD_ROWS) >> 6)];
        const double* uc = d.uvec + d.vptr[c];
        const int* rl = d.rel + d.relptr[c];
        for (int i = ilo + tid; i < ihi; i += MS_THREADS) us[rl[i] - nc] += uc[i];
    }
}

// backward, dots: for 16 columns of B_s, dots[col] = B_s(:, col)' z[rows below]; warp w owns 2 columns
__global__ void __launch_bounds__(MS_THREADS)
mf_bwd_dots_kernel(const MfDesc d, const int2* __restrict__ tasks, const double* __restrict__ x,
                   double* __restrict__ dots) {
    __shared__ double zr[MS_CHUNK];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = tasks[blockIdx.x].x, grp = tasks[blockIdx.x].y;
    const int col0 = d.first[s], nc = d.first[s + 1] - col0, nr = d.nr[s], ld = d.ld[s];
    const int nu = nr - nc;
    const double* B = d.Lv + d.off[s] + d.nb0[s];
    const int* R = d.rows + d.rowptr[s] + nc;
    const int ca = grp * MS_BWD_COLS + warp, cb = ca + 8;
    double acc_a = 0.0, acc_b = 0.0;
    for (int base = 0; base < nu; base += MS_CHUNK) {
        const int len = min(MS_CHUNK, nu - base);
        __syncthreads();
        for (int i = tid; i < len; i += MS_THREADS) zr[i] = x[R[base + i]];
        __syncthreads();
        if (ca < nc) {
            const double* bc = B + base + (long long)ca * ld;
            const double* bd = B + base + (long long)min(cb, nc - 1) * ld;
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            int i = lane;
            for (; i + 96 < len; i += 128) {   // eight loads in flight; same accumulation order as below
                const double p0 = bc[i], q0 = bd[i], p1 = bc[i + 32], q1 = bd[i + 32];
                const double p2 = bc[i + 64], q2 = bd[i + 64], p3 = bc[i + 96], q3 = bd[i + 96];
                a0 = fma(p0, zr[i], a0);
                b0 = fma(q0, zr[i], b0);
                a1 = fma(p1, zr[i + 32], a1);
                b1 = fma(q1, zr[i + 32], b1);
                a0 = fma(p2, zr[i + 64], a0);
                b0 = fma(q2, zr[i + 64], b0);
                a1 = fma(p3, zr[i + 96], a1);
                b1 = fma(q3, zr[i + 96], b1);
            }
            for (; i + 32 < len; i += 64) {
                a0 = fma(bc[i], zr[i], a0);
                b0 = fma(bd[i], zr[i], b0);
                a1 = fma(bc[i + 32], zr[i + 32], a1);
                b1 = fma(bd[i + 32], zr[i + 32], b1);
            }
            if (i < len) {
                a0 = fma(bc[i], zr[i], a0);
                b0 = fma(bd[i], zr[i], b0);
            }
            acc_a += a0 + a1;
            acc_b += b0 + b1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_a += __shfl_xor_sync(0xffffffffu, acc_a, o);
        acc_b += __shfl_xor_sync(0xffffffffu, acc_b, o);
    }
    if (lane == 0) {
        if (ca < nc) dots[col0 + ca] = acc_a;
        if (cb < nc) dots[col0 + cb] = acc_b;
    }
}

// backward, finish: z_s = W_s' (y_s - dots); warp w owns columns w, w+8, ... of W_s
__global__ void __launch_bounds__(MS_THREADS)
mf_bwd_finish_kernel(const MfDesc d, const int* __restrict__ list, double* __restrict__ x,
                     const double* __restrict__ dots) {
    __shared__ double rhs[CH_NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = list[blockIdx.x];
    const int col0 = d.first[s], nc = d.first[s + 1] - col0, nu = d.nr[s] - nc;
    if (tid < CH_NB) rhs[tid] = (tid < nc) ? x[col0 + tid] - (nu > 0 ? dots[col0 + tid] : 0.0) : 0.0;
    __syncthreads();
    const double* Wb = d.W + d.woff[s];
    // z_t = sum_{cc >= t} W(cc, t) rhs[cc]; four columns per step so their loads and shuffle trees overlap
    for (int t0 = warp; t0 < nc; t0 += 32) {
        double a[4], w[4][4];   // nc <= 128: at most four 32-row steps per column
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + 8 * u;
            const double* wc = Wb + (long long)min(t, nc - 1) * nc;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int cc = t + lane + 32 * it;
                w[u][it] = (cc < nc) ? wc[cc] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] = 0.0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int cc = t0 + 8 * u + lane + 32 * it;
                if (cc < nc) a[u] = fma(w[u][it], rhs[cc], a[u]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] += __shfl_xor_sync(0xffffffffu, a[u], o);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (t0 + 8 * u < nc) x[col0 + t0 + 8 * u] = a[u];
        }
    }
}

// multi-GPU: keep only the pieces of the solution this rank is responsible for (its own supernodes;
// rank 0 also keeps the replicated top) so that an all-reduce assembles the whole vector
__global__ void mf_mask_kernel(int nsuper, const int* __restrict__ first, const int* __restrict__ owner,
                               int rank, double* __restrict__ x) {
    const int s = blockIdx.x;
    if (s >= nsuper) return;
    const int o = owner[s];
    if (o == rank || (o < 0 && rank == 0)) return;
    for (int j = first[s] + threadIdx.x; j < first[s + 1]; j += blockDim.x) x[j] = 0.0;
}

__global__ void gather_perm_kernel(int m, const int* __restrict__ perm, const double* __restrict__ src,
                                   double* __restrict__ dst, int inverse) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (!inverse) dst[i] = src[perm[i]];   // y = P b
    else dst[perm[i]] = src[i];            // x = P' y
}


// ------------------------------------------------------------------------------------------------
// host: analysis (upload of the symbolic structure, tensor maps, launch schedule)
// ------------------------------------------------------------------------------------------------
static int sparse_configure(nes_ctx* c) {
    static PerDeviceOnce once;
    int dev;
    if (!once.begin(&dev)) return 0;
    cudaError_t e = cudaFuncSetAttribute(mf_potrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mf_diag_smem(CH_NB));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mf_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mf_tr_smem(CH_NB));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mf_syrk_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SyrkCfg<128>::SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mf_syrk_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SyrkCfg<64>::SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mf_syrk_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(mf_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MF_TI_SMEM);
    once.finish(dev, e == cudaSuccess);
    NES_CUDA(c, e);
    return 0;
}

template <typename T>
static T* up_vec(nes_ctx* c, SparseFactor* sf, const std::vector<T>& v) {
    T* p = static_cast<T*>(dev_alloc(c, (v.size() + 2) * sizeof(T)));
    if (!p) return nullptr;
    sf->owned.push_back(p);
    if (!v.empty() && upload(c, p, v.data(), v.size() * sizeof(T)) != 0) return nullptr;
    return p;
}

template <typename T>
static T* dev_new(nes_ctx* c, SparseFactor* sf, size_t count) {
    T* p = static_cast<T*>(dev_alloc(c, (count + 16) * sizeof(T)));
    if (p) sf->owned.push_back(p);
    return p;
}

int sparse_analyze(nes_ctx* c, nes_matrix* A, nes_factor* L) {
    const MatrixBase* b = A->base;
    const int m = (int)b->m, n = (int)b->n;
    SparseFactor* sf = new SparseFactor();
    L->sparse = sf;
    sf->m = m;
    Symbolic& S = sf->S;
    SymbolicOptions opt;
    opt.nranks = c->nranks;
    opt.nd_leaf = c->nd_leaf;
    char err[256] = {0};
    if (symbolic_analyze(m, n, b->h_colptr.data(), b->h_rowidx.data(), opt, &S, err, sizeof(err)) != 0)
        return fail(c, NES_ERR_INVALID, "%s", err);
    const int ns = S.nsuper;

    // launch schedule: phase A = supernodes of this rank's subtrees, phase B = the replicated top
    std::vector<int> all;
    for (int ph = 0; ph < 2; ++ph) {
        Phase& P = sf->phase[ph];
        std::vector<int> potrf;
        std::vector<int2> trsm;
        std::vector<int4> syrk;
        std::vector<int2> ftail, bdots;
        P.fptr.assign(S.nlevels + 1, 0);
        P.bptr.assign(S.nlevels + 1, 0);
        P.pptr.assign(S.nlevels + 1, 0);
        P.tptr.assign(S.nlevels + 1, 0);
        P.yptr.assign(S.nlevels + 1, 0);
        P.psplit.assign(S.nlevels, 0);
        P.tsplit.assign(S.nlevels, 0);
        for (int l = 0; l < S.nlevels; ++l) {
            for (int wide = 0; wide < 2; ++wide) {  // narrow supernodes first: they launch with less smem
                for (int s = S.lvlptr[l]; s < S.lvlptr[l + 1]; ++s) {
                    const bool mine = (ph == 0) ? (S.owner[s] == c->rank) : (S.owner[s] < 0);
                    const int nc = S.first[s + 1] - S.first[s], nu = S.nr[s] - nc;
                    if (!mine || (nc > 64) != (wide == 1)) continue;
                    potrf.push_back(s);
                    for (int k = 0; k * MF_TR_ROWS < nu; ++k) trsm.push_back(make_int2(s, k));
                    const int tn = (nu + NT_BM - 1) / NT_BM, bn = syrk_bn();
                    for (int bi = 0; bi < tn; ++bi)   // column blocks of width bn up to the end of the diagonal tile
                        for (int cb = 0; cb * bn < std::min(nu, (bi + 1) * NT_BM); ++cb)
                            syrk.push_back(make_int4(s, bi, cb, nc));
                    for (int k = 0; k * MS_FWD_ROWS < nu; ++k) ftail.push_back(make_int2(s, k));
                    if (nu > 0)
                        for (int k = 0; k * MS_BWD_COLS < nc; ++k) bdots.push_back(make_int2(s, k));
                }
                if (wide == 0) {
                    P.psplit[l] = (int)potrf.size();
                    P.tsplit[l] = (int)trsm.size();
                }
            }
            P.pptr[l + 1] = (int)potrf.size();
            P.tptr[l + 1] = (int)trsm.size();
            P.yptr[l + 1] = (int)syrk.size();
            P.fptr[l + 1] = (int)ftail.size();
            P.bptr[l + 1] = (int)bdots.size();
        }
        P.count = (int)potrf.size();
        all.insert(all.end(), potrf.begin(), potrf.end());
        P.d_potrf = up_vec(c, sf, potrf);
        P.d_trsm = up_vec(c, sf, trsm);
        P.d_syrk = up_vec(c, sf, syrk);
        P.d_ftail = up_vec(c, sf, ftail);
        P.d_bdots = up_vec(c, sf, bdots);
        if (!P.d_potrf || !P.d_trsm || !P.d_syrk || !P.d_ftail || !P.d_bdots) return c->status < 0 ? c->status : NES_ERR_OUT_OF_MEMORY;
    }
    sf->nall = (int)all.size();
    sf->d_all = up_vec(c, sf, all);

    std::vector<long long> woff(ns + 1, 0);
    for (int t = 0; t < ns; ++t) {
        const long long nc = S.first[t + 1] - S.first[t];
        woff[t + 1] = woff[t] + nc * nc;
    }
    sf->wsize = woff[ns];

    MfDesc& d = sf->d;
    d.off = up_vec(c, sf, S.off);
    d.uoff = up_vec(c, sf, S.uoff);
    d.woff = up_vec(c, sf, woff);
    d.vptr = up_vec(c, sf, S.vptr);
    d.first = up_vec(c, sf, S.first);
    d.nr = up_vec(c, sf, S.nr);
    d.ld = up_vec(c, sf, S.ld);
    d.ldu = up_vec(c, sf, S.ldu);
    d.rowptr = up_vec(c, sf, S.rowptr);
    d.rows = up_vec(c, sf, S.rows);
    d.childptr = up_vec(c, sf, S.childptr);
    d.child = up_vec(c, sf, S.child);
    d.relptr = up_vec(c, sf, S.relptr);
    d.rel = up_vec(c, sf, S.rel);
    d.cut = up_vec(c, sf, S.cut);
    d.nb0 = up_vec(c, sf, S.nb0);
    d.tbptr = up_vec(c, sf, S.tbptr);
    d.tb = up_vec(c, sf, S.tb);
    sf->d_dots = dev_new<double>(c, sf, (size_t)m);
    sf->d_perm = up_vec(c, sf, S.perm);
    sf->d_ei = up_vec(c, sf, S.ei);
    sf->d_ej = up_vec(c, sf, S.ej);
    sf->d_edest = up_vec(c, sf, S.edest);
    sf->d_owner = up_vec(c, sf, S.owner);
    d.Lv = dev_new<double>(c, sf, (size_t)S.lsize);
    d.U = dev_new<double>(c, sf, (size_t)S.usize);
    d.dinv = dev_new<double>(c, sf, (size_t)m);
    d.W = dev_new<double>(c, sf, (size_t)sf->wsize);
    d.uvec = dev_new<double>(c, sf, (size_t)S.vsize);
    sf->d_x = dev_new<double>(c, sf, (size_t)m);
    sf->d_info = dev_new<int>(c, sf, 4);
    d.info = sf->d_info;
    L->d_rhs = static_cast<double*>(dev_alloc(c, (size_t)(m + 16) * sizeof(double)));
    const bool ok = d.off && d.uoff && d.woff && d.vptr && d.first && d.nr && d.ld && d.ldu && d.rowptr && d.rows &&
                    d.childptr && d.child && d.relptr && d.rel && d.cut && d.nb0 && d.tbptr && d.tb && sf->d_dots && sf->d_perm && sf->d_ei && sf->d_ej &&
                    sf->d_edest && sf->d_owner && sf->d_all && d.Lv && d.U && d.dinv && d.W && d.uvec && sf->d_x &&
                    sf->d_info && L->d_rhs;
    if (!ok) return c->status < 0 ? c->status : NES_ERR_OUT_OF_MEMORY;
    NES_CUDA(c, cudaMemsetAsync(d.dinv, 0, (size_t)m * sizeof(double), c->stream));
    NES_CUDA(c, cudaMemsetAsync(d.uvec, 0, (size_t)(S.vsize + 16) * sizeof(double), c->stream));

    // one tensor map per supernode block (132 x 32 boxes, OOB rows / columns read as zero)
    {
        std::vector<CUtensorMap> maps(2 * (size_t)ns);  // [0, ns): 132-row boxes, [ns, 2 ns): 68-row boxes
        for (int s = 0; s < ns; ++s) {
            const int nc = S.first[s + 1] - S.first[s];
            if (S.nr[s] == nc) {
                memset(&maps[s], 0, sizeof(CUtensorMap));
                memset(&maps[ns + s], 0, sizeof(CUtensorMap));
                continue;
            }
            if (make_operand_map(&maps[s], d.Lv + S.off[s], S.nb0[s] + S.nr[s] - nc, nc, S.ld[s]) != 0 ||
                make_operand_map(&maps[ns + s], d.Lv + S.off[s], S.nb0[s] + S.nr[s] - nc, nc, S.ld[s], 68, NT_BK) != 0)
                return fail(c, NES_ERR_CUDA, "cuTensorMapEncodeTiled failed for supernode %d (%d x %d)", s, S.nr[s], nc);
        }
        CUtensorMap* dm = static_cast<CUtensorMap*>(dev_alloc(c, (2 * (size_t)ns + 1) * sizeof(CUtensorMap)));
        if (!dm) return c->status;
        sf->owned.push_back(dm);
        NES_TRY(upload(c, dm, maps.data(), 2 * (size_t)ns * sizeof(CUtensorMap)));
        d.maps = dm;
        d.maps68 = dm + ns;
    }
    c->anz = (double)S.anz;
    c->aatfl = S.aatfl;
    c->lnz = S.lnz;
    c->fl = S.fl;
    c->status = 0;
    return 0;
}

void sparse_free(nes_ctx* c, nes_factor* L) {
    SparseFactor* sf = L->sparse;
    if (!sf) return;
    for (auto& g : sf->g_factor)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& g : sf->g_solve)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (void* p : sf->owned) dev_free(c, p);
    delete sf;
    L->sparse = nullptr;
}

// ------------------------------------------------------------------------------------------------
// host: numeric factorization
// ------------------------------------------------------------------------------------------------
// Level loop of one phase on the library's stream.  The low-priority side stream computes the inverses
// W_s of the diagonal blocks (needed by the solves only) while the narrow top of the tree leaves most
// SMs idle; they used to be one 0.7 ms launch at the end.
static int run_factor_phase(nes_ctx* c, SparseFactor* sf, int ph) {
    const Phase& P = sf->phase[ph];
    const Symbolic& S = sf->S;
    int trtri_done = 0;  // supernodes [0, trtri_done) of this phase's list have their inverse enqueued
    for (int l = 0; l < S.nlevels; ++l) {
        if (P.pptr[l + 1] == P.pptr[l]) continue;
        // Narrow (nc <= 64) and wide supernodes launch separately (different shared-memory footprints) and
        // do not depend on each other: when a level has both, the narrow chain potrf -> trsm runs on the
        // second stream beside the wide chain and joins before the SYRK (the level's critical path is one
        // potrf + one trsm instead of two of each; this is what bounds the narrow top of the tree).
        static const bool fork_off = getenv("NES_SPARSE_NO_FORK") != nullptr;  // comparison runs
        const bool fork = c->stream_b && !fork_off && P.psplit[l] > P.pptr[l] && P.pptr[l + 1] > P.psplit[l] &&
                          !sparse_sync_debug();
        if (fork) {
            NES_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
            NES_CUDA(c, cudaStreamWaitEvent(c->stream_b, c->ev_fork, 0));
        }
        for (int wide = 0; wide < 2; ++wide) {
            cudaStream_t st = (fork && !wide) ? c->stream_b : c->stream;
            const int ncmax = wide ? CH_NB : 64;
            const int a = wide ? P.psplit[l] : P.pptr[l], b = wide ? P.pptr[l + 1] : P.psplit[l];
            if (b > a) {
                mf_potrf_kernel<<<b - a, 256, mf_diag_smem(ncmax), st>>>(sf->d, P.d_potrf + a, ncmax);
                MF_LAUNCHED(c, "mf_potrf_kernel");
            }
            const int ta = wide ? P.tsplit[l] : P.tptr[l], tb = wide ? P.tptr[l + 1] : P.tsplit[l];
            if (tb > ta) {
                mf_trsm_kernel<<<tb - ta, 256, mf_tr_smem(ncmax), st>>>(sf->d, P.d_trsm + ta, ncmax);
                MF_LAUNCHED(c, "mf_trsm_kernel");
            }
        }
        if (fork) {
            NES_CUDA(c, cudaEventRecord(c->ev_join, c->stream_b));
            NES_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        }
        // block inverses: held back while the levels are wide (they would take SMs from the factorization),
        // released in one go when the tree narrows to <= 32 supernodes per level and the SMs are mostly idle
        const int width = P.pptr[l + 1] - P.pptr[l];
        if (width <= 32 || l + 1 == S.nlevels || P.pptr[l + 1] == P.count) {
            NES_CUDA(c, cudaEventRecord(c->ev_panel, c->stream));
            NES_CUDA(c, cudaStreamWaitEvent(c->stream_aux, c->ev_panel, 0));
            const int from = trtri_done;
            mf_trtri_kernel<<<P.pptr[l + 1] - from, TRTRI_THREADS, MF_TI_SMEM, c->stream_aux>>>(sf->d, P.d_potrf + from);
            NES_CHECK_LAUNCH(c);
            trtri_done = P.pptr[l + 1];
        }
        const int ny = P.yptr[l + 1] - P.yptr[l];
        if (ny > 0) {
            if (syrk_bn() == 128) {
                const int grid = ny < c->num_sms ? ny : c->num_sms;
                mf_syrk_kernel<128><<<grid, SyrkCfg<128>::THREADS, SyrkCfg<128>::SMEM, c->stream>>>(
                    sf->d, P.d_syrk + P.yptr[l], ny);
            } else {
                const int grid = ny < 2 * c->num_sms ? ny : 2 * c->num_sms;
                mf_syrk_kernel<64><<<grid, SyrkCfg<64>::THREADS, SyrkCfg<64>::SMEM, c->stream>>>(
                    sf->d, P.d_syrk + P.yptr[l], ny);
            }
            MF_LAUNCHED(c, "mf_syrk_kernel");
        }
    }
    return 0;
}

__global__ void info_first_minor_kernel(int* info) {
    if (threadIdx.x == 0 && info[0] == 0) info[1] = 0x7fffffff;
}

static int enqueue_factorization(nes_ctx* c, SparseFactor* sf, const MatrixBase* b, const double* d_theta) {
    const Symbolic& S = sf->S;
    NES_CUDA(c, cudaMemsetAsync(sf->d.Lv, 0, (size_t)S.lsize * sizeof(double), c->stream));
    NES_CUDA(c, cudaMemsetAsync(sf->d_info, 0, 2 * sizeof(int), c->stream));
    // (assembling the levels above the bottom one on the side stream was measured: the low-priority
    // kernel starves behind the bottom level and the factorization then waits for it, +2.2 ms)
    if (S.anz > 0) {
        sparse_assemble_kernel<<<(unsigned)((S.anz + 255) / 256), 256, 0, c->stream>>>(
            S.anz, sf->d_ei, sf->d_ej, sf->d_edest, b->d_rowptr, b->d_colidx, b->d_csr_val, d_theta, sf->d.Lv);
        MF_LAUNCHED(c, "sparse_assemble_kernel");
    }
    NES_TRY(run_factor_phase(c, sf, 0));
    return 0;
}

int sparse_factorize(nes_ctx* c, nes_matrix* A, nes_factor* L) {
    SparseFactor* sf = L->sparse;
    if (!sf) return fail(c, NES_ERR_INVALID, "sparse factor was not analyzed");
    NES_TRY(sparse_configure(c));
    const MatrixBase* b = A->base;
    const Symbolic& S = sf->S;
    L->factorized = 0;
    sf->d.dbound = c->dbound;
    if (c->nranks == 1) {
        StageTimer t(c, NES_STAGE_FACTOR);
        const double* d_theta = A->d_theta;
        SparseFactor::GraphCache* gc = nullptr;
        for (auto& cand : sf->g_factor)
            if (cand.exec && cand.k0 == d_theta && cand.k1 == b->d_csr_val && cand.k2 == c->dbound) gc = &cand;
        if (!gc) {
            gc = &sf->g_factor[sf->g_factor_next];
            sf->g_factor_next = (sf->g_factor_next + 1) % 4;
        }
        NES_TRY(run_captured(c, sf, *gc, d_theta, b->d_csr_val, c->dbound, [&]() -> int {
            NES_TRY(enqueue_factorization(c, sf, b, d_theta));
            NES_CUDA(c, cudaEventRecord(c->ev_aux, c->stream_aux));  // join: the block inverses are complete
            NES_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_aux, 0));
            return 0;
        }));
    } else {
        StageTimer t(c, NES_STAGE_FACTOR);
        NES_TRY(enqueue_factorization(c, sf, b, A->d_theta));
        if (c->nranks > 1) {
            // publish the update matrices of my subtree roots: contiguous per owner, one broadcast each
            for (int q = 0; q < c->nranks; ++q) {
                const long long cnt = S.xu_off[q + 1] - S.xu_off[q];
                if (cnt > 0) NES_TRY(dist_broadcast(c, sf->d.U + S.xu_off[q], (size_t)cnt, q));
            }
        }
        NES_TRY(run_factor_phase(c, sf, 1));
        NES_CUDA(c, cudaEventRecord(c->ev_aux, c->stream_aux));  // join: the block inverses are complete
        NES_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_aux, 0));
        if (c->nranks > 1) {  // a failed pivot is seen by one rank only: agree on {status, first minor}
            info_first_minor_kernel<<<1, 32, 0, c->stream>>>(sf->d_info);
            NES_CHECK_LAUNCH(c);
            NES_TRY(dist_allreduce_int(c, sf->d_info, 1, 1));
            NES_TRY(dist_allreduce_int(c, sf->d_info + 1, 1, 0));
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
    return 0;
}

// ------------------------------------------------------------------------------------------------
// host: solves
// ------------------------------------------------------------------------------------------------
static int run_solve_phase(nes_ctx* c, SparseFactor* sf, int ph, bool backward) {
    const Phase& P = sf->phase[ph];
    const int nl = sf->S.nlevels;
    for (int k = 0; k < nl; ++k) {
        const int l = backward ? nl - 1 - k : k;
        const int np = P.pptr[l + 1] - P.pptr[l];
        if (np == 0) continue;
        if (!backward) {
            mf_fwd_head_kernel<<<np, MS_THREADS, 0, c->stream>>>(sf->d, P.d_potrf + P.pptr[l], sf->d_x);
            MF_LAUNCHED(c, "mf_fwd_head_kernel");
            const int nt = P.fptr[l + 1] - P.fptr[l];
            if (nt > 0) {
                mf_fwd_tail_kernel<<<nt, MS_THREADS, 0, c->stream>>>(sf->d, P.d_ftail + P.fptr[l], sf->d_x);
                MF_LAUNCHED(c, "mf_fwd_tail_kernel");
            }
        } else {
            const int nt = P.bptr[l + 1] - P.bptr[l];
            if (nt > 0) {
                mf_bwd_dots_kernel<<<nt, MS_THREADS, 0, c->stream>>>(sf->d, P.d_bdots + P.bptr[l], sf->d_x, sf->d_dots);
                MF_LAUNCHED(c, "mf_bwd_dots_kernel");
            }
            mf_bwd_finish_kernel<<<np, MS_THREADS, 0, c->stream>>>(sf->d, P.d_potrf + P.pptr[l], sf->d_x, sf->d_dots);
            MF_LAUNCHED(c, "mf_bwd_finish_kernel");
        }
    }
    return 0;
}

int sparse_solve_inplace(nes_ctx* c, nes_factor* L, double* d_x) {
    SparseFactor* sf = L->sparse;
    StageTimer t(c, NES_STAGE_SOLVE);
    const Symbolic& S = sf->S;
    const int m = sf->m;
    auto enqueue = [&]() -> int {
        gather_perm_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(m, sf->d_perm, d_x, sf->d_x, 0);
        NES_CHECK_LAUNCH(c);
        NES_TRY(run_solve_phase(c, sf, 0, false));
        if (c->nranks > 1) {
            for (int q = 0; q < c->nranks; ++q) {
                const long long cnt = S.xv_off[q + 1] - S.xv_off[q];
                if (cnt > 0) NES_TRY(dist_broadcast(c, sf->d.uvec + S.xv_off[q], (size_t)cnt, q));
            }
        }
        NES_TRY(run_solve_phase(c, sf, 1, false));
        NES_TRY(run_solve_phase(c, sf, 1, true));
        NES_TRY(run_solve_phase(c, sf, 0, true));
        if (c->nranks > 1) {
            mf_mask_kernel<<<S.nsuper, 128, 0, c->stream>>>(S.nsuper, sf->d.first, sf->d_owner, c->rank, sf->d_x);
            NES_CHECK_LAUNCH(c);
            NES_TRY(dist_allreduce_sum(c, sf->d_x, (size_t)m));
        }
        gather_perm_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(m, sf->d_perm, sf->d_x, d_x, 1);
        NES_CHECK_LAUNCH(c);
        return 0;
    };
    SparseFactor::GraphCache* g = nullptr;
    for (auto& cand : sf->g_solve)
        if (cand.exec && cand.k0 == d_x) g = &cand;
    if (!g) {
        g = &sf->g_solve[sf->g_solve_next];
        sf->g_solve_next = (sf->g_solve_next + 1) % 4;
    }
    return run_captured(c, sf, *g, d_x, nullptr, 0.0, enqueue);
}

// expand the supernodal factor to a dense lower-triangular matrix (testing) + permutation
int sparse_factor_to_dense(nes_ctx* c, nes_factor* L, double* Lout, size_t ldo, int* perm_out) {
    SparseFactor* sf = L->sparse;
    const Symbolic& S = sf->S;
    if (c->nranks > 1) {  // the blocks of a subtree live on its owner only: collect them (testing path)
        for (int s = 0; s < S.nsuper; ++s)
            if (S.owner[s] >= 0)
                NES_TRY(dist_broadcast(c, sf->d.Lv + S.off[s], (size_t)(S.off[s + 1] - S.off[s]), S.owner[s]));
    }
    std::vector<double> h((size_t)S.lsize);
    NES_TRY(download(c, h.data(), sf->d.Lv, h.size() * sizeof(double)));
    for (size_t j = 0; j < (size_t)sf->m; ++j)
        for (size_t i = 0; i < (size_t)sf->m; ++i) Lout[i + j * ldo] = 0.0;
    for (int s = 0; s < S.nsuper; ++s) {
        const int col0 = S.first[s], nc = S.first[s + 1] - col0, nr = S.nr[s];
        const int* R = S.rows.data() + S.rowptr[s];
        for (int cc = 0; cc < nc; ++cc)
            for (int r = cc; r < nr; ++r)
                Lout[(size_t)R[r] + (size_t)(col0 + cc) * ldo] =
                    h[(size_t)S.off[s] + (r < nc ? r : r - nc + S.nb0[s]) + (size_t)cc * S.ld[s]];
    }
    if (perm_out)
        for (int i = 0; i < sf->m; ++i) perm_out[i] = S.perm[i];
    return 0;
}

}  // namespace nes
