// FP64 tensor-core "NT" contraction for column-major operands, the workhorse of the engine:
//
//     C(MxN) = alpha * X[rowA0.., k0..k0+K) * diag(scale) * Y[rowB0.., k0..k0+K)^T + beta * C
//
// X, Y are column-major (rows contiguous), so a BMxBK operand tile is BK runs of BM contiguous
// doubles: exactly what one 2D TMA box delivers.  Used for
//   * K1  formation  M = A diag(theta) A^T      (scale = theta, lower tiles only, X = Y = A)
//         replaces matlisp m* by an n x n diagonal + CHOLMOD's A*A' assembly
//         (reference newton-solve.lisp:112-116, sparse-cholesky.lisp:409-419)
//   * K2  Cholesky trailing update  C -= P P^T  (alpha=-1, beta=1, lower tiles only, X = Y = panel)
//         replaces dsyrk/dgemm inside cholmod_factorize (sparse-cholesky.lisp:419)
//   * general rectangular updates (look-ahead column, supernode updates).
//
// Structure (one persistent CTA per SM, static tile round-robin):
//   thread 0      : also the TMA producer.  It issues cp.async.bulk.tensor for the X tile, the Y tile
//                   (skipped on diagonal tiles of a symmetric product) and a 1D bulk copy of the BK
//                   scale values, all completing on the stage's "full" mbarrier.  (A dedicated
//                   producer warp would make 9 warps; registers are allocated per 4 warps, which
//                   caps the consumers at 168 registers and spills the accumulators.)
//   warps 0..7    : DMMA consumers, 2 (M) x 4 (N); warp tile 64x32 = 8x4 m8n8k4 accumulators
//                   (128 accumulator registers per thread).  Each k4 step: 8+4 LDS.64 -> 32 DMMA.
//
// Shared-memory layout: the TMA box is 132 rows x BK columns although only 128 rows are used.  The
// dense box lands as [k][132]; with a row pitch of 132 doubles (== 4 mod 16) the fragment loads of
// m8n8k4 (lane = 4*g + t reads row g, column t) hit 16 distinct 8-byte bank pairs per half-warp, so
// they are conflict-free without swizzling.  Cost: 3% extra L2->smem traffic.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx_util.cuh"

namespace nes {

constexpr int NT_BM = 128;
constexpr int NT_BN = 128;
#ifndef NT_BK_OVERRIDE
#define NT_BK_OVERRIDE 32
#endif
constexpr int NT_BK = NT_BK_OVERRIDE;
constexpr int NT_PITCH = 132;  // smem row pitch in doubles (TMA box rows)
#ifndef NT_STAGES_OVERRIDE
#define NT_STAGES_OVERRIDE 3
#endif
constexpr int NT_STAGES = NT_STAGES_OVERRIDE;
constexpr int NT_CONSUMER_WARPS = 8;
constexpr int NT_THREADS = NT_CONSUMER_WARPS * 32;
constexpr int NT_TILE_BYTES = NT_PITCH * NT_BK * 8;            // 16896
constexpr int NT_STAGE_BYTES = 2 * NT_TILE_BYTES + NT_BK * 8;  // + scale slab
constexpr int NT_SMEM_BYTES = NT_STAGES * NT_STAGE_BYTES + 2 * NT_STAGES * 8 + 16 + 128;

struct NtArgs {
    double* C;            // origin of the output region (column-major)
    long long ldc;        // leading dimension of C in doubles
    int M, N;             // region extent
    int rowA0, rowB0;     // operand row origins (coordinates in the tensor maps)
    int k0, K;            // contraction range: operand columns [k0, k0+K)
    const double* scale;  // per-column scale, indexed by absolute column; padded to a multiple of BK
    double alpha, beta;
    int lower;            // only tiles with bj <= bi (region must be square, same row origin)
    int same_operand;     // X == Y and rowA0 == rowB0: diagonal tiles load one operand tile
    int ntiles;
    int tiles_n;
    // tail balancing: the last `split_r` tiles (the partial wave of the static round-robin) are cut
    // into `split_s` k-ranges each; partial tiles go to split_ws and the last CTA to arrive sums them
    // in k-range order (fixed order => bitwise reproducible, no floating-point atomics).
    int split_r, split_s;
    int split_all;        // every tile is split (launches with fewer tiles than SMs); split_r is then ntiles
    double* split_ws;     // split_r * split_s * 128*128 doubles
    int* split_counters;  // split_r ints, zero on entry, reset to zero by the last arriver
    // explicit tile list (multi-GPU ownership): tile t is (tile_list[t].x, tile_list[t].y) in absolute
    // 128-row / 128-column units of the region; overrides the lower/rectangular enumeration
    const int2* tile_list;
    // batched mode (many small independent problems stacked along the rows): tile t belongs to problem
    // t / batch_tiles; operand rows, C rows and the scale vector are offset per problem
    int batch_tiles;            // tiles per problem (0 = not batched)
    int batch_rows;             // row stride between problems in the operand / C matrices
    int batch_scale_stride;     // stride between the problems' scale vectors
};

// Tile order.  A wave of the persistent grid is ~148 consecutive tiles and only tiles that run at the
// same time share operand rows through L2 (one 128-row block of A is 8*128*n bytes, far more than L2
// holds for all blocks), so consecutive tiles should form a squarish patch: the lower triangle is cut
// into bands of NT_BAND tile rows, each band walked column by column (a wave then touches about
// NT_BAND + 148/NT_BAND row blocks instead of ~64).
#ifndef NT_BAND_ROWS
#define NT_BAND_ROWS 12
#endif
constexpr int NT_BAND = NT_BAND_ROWS;

__device__ __forceinline__ void nt_tile_coords(const NtArgs& p, int t, int& bi, int& bj) {
    if (p.tile_list) {
        const int2 v = p.tile_list[t];
        bi = v.x;
        bj = v.y;
    } else if (p.lower) {
        // band r holds tile rows [r*H, r*H+h); tiles before band r: (rH)(rH+1)/2
        const int H = NT_BAND;
        int r = static_cast<int>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5) / H;
        while ((long long)((r + 1) * H) * ((r + 1) * H + 1) / 2 <= t) ++r;
        while ((long long)(r * H) * (r * H + 1) / 2 > t) --r;
        const int r0 = r * H;
        const int h = min(H, p.tiles_n - r0);
        int q = t - static_cast<int>((long long)r0 * (r0 + 1) / 2);
        if (q < r0 * h) {  // full-height columns left of the band's diagonal block
            bj = q / h;
            bi = r0 + q - bj * h;
        } else {           // the h x h lower triangle on the diagonal, column by column
            q -= r0 * h;
            int cj = 0;
            while (q >= h - cj) {
                q -= h - cj;
                ++cj;
            }
            bj = r0 + cj;
            bi = bj + q;
        }
    } else {
        bi = t / p.tiles_n;
        bj = t - bi * p.tiles_n;
    }
}

// One unit of work for a CTA: a tile and a range of k-chunks of it.
struct NtWork {
    int tile, bi, bj;
    int brow;  // row offset of the problem this tile belongs to (batched mode), else 0
    int kc_begin, kc_end;
    int split;  // -1: whole tile, else k-range index of a split tail tile
};

// Items 0..nfull-1 are whole tiles, nfull.. are (tail tile, k-range) pairs; CTA b takes b, b+G, ...
__device__ __forceinline__ bool nt_get_work(const NtArgs& p, int kchunks, int item, NtWork& w) {
    const int nfull = p.ntiles - p.split_r;
    if (item < nfull) {
        w.tile = item;
        w.kc_begin = 0;
        w.kc_end = kchunks;
        w.split = -1;
    } else {
        const int q = item - nfull;
        if (q >= p.split_r * p.split_s) return false;
        w.tile = nfull + q / p.split_s;
        w.split = q - (q / p.split_s) * p.split_s;
        const int per = (kchunks + p.split_s - 1) / p.split_s;
        w.kc_begin = min(kchunks, w.split * per);
        w.kc_end = min(kchunks, w.kc_begin + per);
    }
    if (p.batch_tiles > 0) {
        const int bb = w.tile / p.batch_tiles;
        nt_tile_coords(p, w.tile - bb * p.batch_tiles, w.bi, w.bj);
        w.brow = bb * p.batch_rows;
    } else {
        nt_tile_coords(p, w.tile, w.bi, w.bj);
        w.brow = 0;
    }
    return true;
}

template <bool kHasScale>
__global__ void __launch_bounds__(NT_THREADS, 1)
dmma_nt_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
               const NtArgs p) {
    // The kernel has no static __shared__, so the dynamic window starts at the CTA's shared base
    // (128B-aligned for the TMA destinations) and `smem` stays a shared-typed pointer: the fragment
    // loads compile to LDS instead of generic LD.
#ifndef NT_SHARED_TYPED_SMEM
    // measured on B200 (m=8192, n=16384): generic addressing of the staging buffers schedules better
    // than a shared-typed pointer (LDS): 34.80 vs 35.33 ms; BK=32 x 3 stages beats 16 x 5 (34.24 ms)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
#else
    extern __shared__ __align__(128) uint8_t smem[];
#endif
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NT_STAGES * NT_STAGE_BYTES);
    uint64_t* empty = full + NT_STAGES;
    int& s_last = *reinterpret_cast<int*>(empty + NT_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool is_producer = (threadIdx.x == 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NT_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NT_CONSUMER_WARPS);
        }
        fence_mbar_init();
        prefetch_tmap(&mapX);
        prefetch_tmap(&mapY);
    }
    __syncthreads();

    const int kchunks = (p.K + NT_BK - 1) / NT_BK;

    // The TMA producer is folded into thread 0: it refills the stage released one chunk ago, so its
    // empty-barrier wait is normally already satisfied (prefetch distance STAGES-1 chunks).
    NtWork pw;            // producer's current work item
    int p_item = blockIdx.x;
    int p_kc = 0;
    uint32_t p_it = 0;
    bool p_valid = false;
    if (is_producer) {
        p_valid = nt_get_work(p, kchunks, p_item, pw);
        while (p_valid && pw.kc_begin >= pw.kc_end) {  // empty k-range (never for whole tiles)
            p_item += gridDim.x;
            p_valid = nt_get_work(p, kchunks, p_item, pw);
        }
        if (p_valid) p_kc = pw.kc_begin;
    }
    auto produce = [&]() {
        if (!p_valid) return;
        const int s = p_it % NT_STAGES;
        const uint32_t ph = (p_it / NT_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const bool diag = p.same_operand && (pw.bi == pw.bj);
        const uint32_t bytes =
            (diag ? NT_TILE_BYTES : 2 * NT_TILE_BYTES) + (kHasScale ? NT_BK * 8 : 0);
        uint8_t* st = smem + s * NT_STAGE_BYTES;
        mbar_expect_tx(&full[s], bytes);
        const int kk = p.k0 + p_kc * NT_BK;
        tma_load_2d(st, &mapX, p.rowA0 + pw.brow + pw.bi * NT_BM, kk, &full[s]);
        if (!diag)
            tma_load_2d(st + NT_TILE_BYTES, &mapY, p.rowB0 + pw.brow + pw.bj * NT_BN, kk, &full[s]);
        if (kHasScale) {
            const double* sc = p.scale + kk;
            if (p.batch_tiles > 0) sc += (long long)(pw.brow / p.batch_rows) * p.batch_scale_stride;
            bulk_load_1d(st + 2 * NT_TILE_BYTES, sc, NT_BK * 8, &full[s]);
        }
        ++p_it;
        if (++p_kc >= pw.kc_end) {
            do {
                p_item += gridDim.x;
                p_valid = nt_get_work(p, kchunks, p_item, pw);
            } while (p_valid && pw.kc_begin >= pw.kc_end);
            if (p_valid) p_kc = pw.kc_begin;
        }
    };
    if (is_producer) {
        for (int i = 0; i < NT_STAGES - 1; ++i) produce();
    }

    // ---------------- DMMA consumers ----------------
    const int g = lane >> 2;
    const int t4 = lane & 3;
    const int wm = warp >> 2;  // 0..1  -> 64-row slab
    const int wn = warp & 3;   // 0..3  -> 32-col slab
    const int a_off = t4 * NT_PITCH + wm * 64 + g;
    const int b_off = t4 * NT_PITCH + wn * 32 + g;

    uint32_t it = 0;
    NtWork w;
    for (int item = blockIdx.x; nt_get_work(p, kchunks, item, w); item += gridDim.x) {
        const int bi = w.bi, bj = w.bj;
        const bool diag = p.same_operand && (bi == bj);
        const int row_base = bi * NT_BM + wm * 64 + g;
        const int col_base = bj * NT_BN + wn * 32 + 2 * t4;
        double* const Cb = p.C + w.brow;  // batched: the problem's rows inside the stacked C

        // warm L2 with the C tile while the mainloop runs (beta path reads it in the epilogue)
        if (p.beta != 0.0 && g == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = col_base + j * 8 + c;
                    if (col < p.N) {
                        const double* cp = Cb + (long long)col * p.ldc;
#pragma unroll
                        for (int i = 0; i < 8; i += 2)
                            if (row_base + i * 8 < p.M) prefetch_l2(cp + row_base + i * 8);
                    }
                }
        }

        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kc = w.kc_begin; kc < w.kc_end; ++kc, ++it) {
            if (is_producer) produce();
            __syncwarp();
            const int s = it % NT_STAGES;
            const uint32_t ph = (it / NT_STAGES) & 1;
            mbar_wait(&full[s], ph);
            const double* sA = reinterpret_cast<const double*>(smem + s * NT_STAGE_BYTES);
            const double* sB = diag ? sA : sA + NT_PITCH * NT_BK;
            const double* sS = sA + 2 * NT_PITCH * NT_BK;
            const double* ap = sA + a_off;
            const double* bp = sB + b_off;
#pragma unroll
            for (int ks = 0; ks < NT_BK / 4; ++ks) {
                double af[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) af[i] = ap[ks * 4 * NT_PITCH + i * 8];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = bp[ks * 4 * NT_PITCH + j * 8];
                if (kHasScale) {
                    const double th = sS[ks * 4 + t4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bf[j] *= th;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            // Plain arrive.  The hazard mbar_arrive_after (ptx_util.cuh) cures -- a fragment load that is still in
            // the load/store unit when the stage is refilled -- needs a backlog of global traffic ahead of the
            // load.  With ONE CTA per SM the only source is this CTA's own previous epilogue, which is fenced off
            // below; the dependency-carrying arrive costs this kernel 2.7% (all chunks) to 4% (first chunks only:
            // the duplicated loop body no longer fits the instruction cache), measured, and is used by the
            // kernels that keep two CTAs on an SM (dmma_nt64, mf_syrk), where the hazard was observed.
            if (lane == 0) mbar_arrive(&empty[s]);
        }

        if (w.split >= 0) {
            // ---- split tail tile: park the partial tile, last arriver reduces in k-range order ----
            const int slot = w.tile - (p.ntiles - p.split_r);
            double* ws = p.split_ws + ((size_t)slot * p.split_s + w.split) * (NT_BM * NT_BN);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        __stcg(ws + ((i * 4 + j) * 2 + c) * NT_THREADS + threadIdx.x, acc[i][j][c]);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                const int old = atomicAdd(p.split_counters + slot, 1);
                s_last = (old == p.split_s - 1);
                if (s_last) p.split_counters[slot] = 0;
            }
            __syncthreads();
            if (!s_last) continue;
            __threadfence();
            // k-range order for every element (fixed => reproducible); 16 independent loads per batch so
            // the reduction is not a chain of dependent L2 round trips
            const double* base = p.split_ws + (size_t)slot * p.split_s * (NT_BM * NT_BN) + threadIdx.x;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int sp = 0; sp < p.split_s; ++sp) {
                const double* q = base + (size_t)sp * (NT_BM * NT_BN);
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    double v[2][4][2];
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                v[ii][j][c] = __ldcg(q + (((i + ii) * 4 + j) * 2 + c) * NT_THREADS);
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int c = 0; c < 2; ++c) acc[i + ii][j][c] += v[ii][j][c];
                }
            }
        }

        // epilogue: registers -> global (column-major).  Lane holds rows g (+8i), cols 2*t4, 2*t4+1.
        // The beta path loads a whole column pair (16 values) before storing, so the loads overlap.
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double old[2][8];
            if (p.beta != 0.0) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = col_base + j * 8 + c;
                    const double* cp = Cb + (long long)col * p.ldc;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row_base + i * 8;
                        old[c][i] = (col < p.N && row < p.M) ? __ldcg(cp + row) : 0.0;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int col = col_base + j * 8 + c;
                if (col < p.N) {
                    double* cp = Cb + (long long)col * p.ldc;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row_base + i * 8;
                        if (row < p.M) {
                            double v = p.alpha * acc[i][j][c];
                            if (p.beta != 0.0) v = fma(p.beta, old[c][i], v);
                            cp[row] = v;
                        }
                    }
                }
            }
        }
        // persistent grids: do not start the next tile's fragment loads behind this tile's 128 KB of stores
        // (see the note at the stage release above)
        __threadfence_block();
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
                cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// Tensor map over a column-major rows x cols matrix of doubles (ld even, base 16B aligned) with
// the 132 x 16 operand box of dmma_nt_kernel.  Returns 0 on success.
inline int make_operand_map(CUtensorMap* map, const double* base, long long rows, long long cols,
                            long long ld, int box_rows = NT_PITCH, int box_cols = NT_BK) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return -1;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(cols)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 8};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_cols)};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char* e = getenv("NES_TMA_L2PROMO")) {  // debugging: 0 none, 64, 128, 256
        const int v = atoi(e);
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                       : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                  : (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

inline cudaError_t nt_configure() {
    static PerDeviceOnce once;  // per translation unit (each has its own copy of the kernel) and per device
    int dev;
    if (!once.begin(&dev)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(dmma_nt_kernel<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, NT_SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(dmma_nt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 NT_SMEM_BYTES);
    once.finish(dev, e == cudaSuccess);
    return e;
}

// Tail-balancing plan for a launch with `ntiles` equal-cost tiles of `kchunks` k-chunks on `grid`
// CTAs: returns the number of tail tiles to split (0 = none) and the k-ranges per tile.
inline void nt_plan_split(int ntiles, int kchunks, int grid, int* split_r, int* split_s) {
    *split_r = *split_s = 0;
    if (ntiles <= grid) return;
    const int r = ntiles % grid;
    if (r == 0) return;
    int s = grid / r;
    if (s > kchunks / 8) s = kchunks / 8;  // keep every k-range at least 8 chunks long
    if (s < 2) return;
    *split_r = r;
    *split_s = s;
}

// Launch on `stream` with at most `max_ctas` persistent CTAs (normally the SM count).
inline cudaError_t nt_launch(const CUtensorMap& mapX, const CUtensorMap& mapY, NtArgs a, int max_ctas,
                             cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0) return cudaSuccess;
    cudaError_t e = nt_configure();
    if (e != cudaSuccess) return e;
    const int tm = (a.M + NT_BM - 1) / NT_BM;
    const int tn = (a.N + NT_BN - 1) / NT_BN;
    a.tiles_n = tn;
    if (a.batch_tiles > 0) {
        // caller sets ntiles = problems * batch_tiles; batch_tiles must match the per-problem count
        a.batch_tiles = a.lower ? tm * (tm + 1) / 2 : tm * tn;
    } else if (!a.tile_list) {
        a.ntiles = a.lower ? tm * (tm + 1) / 2 : tm * tn;
    }
    if (a.ntiles <= 0) return cudaSuccess;
    int grid = a.ntiles < max_ctas ? a.ntiles : max_ctas;
    if (a.split_ws == nullptr || a.split_counters == nullptr || a.split_s < 2)
        a.split_r = a.split_s = a.split_all = 0;
    if (a.split_all) {
        a.split_r = a.ntiles;
        const long long items = (long long)a.ntiles * a.split_s;
        grid = items < max_ctas ? (int)items : max_ctas;
    } else if (a.split_r > 0) {
        // caller proposes; validate against this grid
        const int r = a.ntiles % grid;
        if (r != a.split_r || a.split_s < 2 || a.split_r * a.split_s > grid) a.split_r = a.split_s = 0;
    }
    if (a.scale)
        dmma_nt_kernel<true><<<grid, NT_THREADS, NT_SMEM_BYTES, stream>>>(mapX, mapY, a);
    else
        dmma_nt_kernel<false><<<grid, NT_THREADS, NT_SMEM_BYTES, stream>>>(mapX, mapY, a);
    return cudaGetLastError();
}

}  // namespace nes
