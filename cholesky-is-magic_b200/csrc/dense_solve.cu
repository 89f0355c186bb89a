// K5: forward / backward triangular solves with the dense Cholesky factor, one right-hand side.
// Replaces cholmod_solve(CHOLMOD_A, L, b) / cholmod_solve2 (sparse-cholesky.lisp:422, 515, 546).
// HBM-bound: each sweep reads tril(L) once (4 m^2 bytes).  Blocked by 128:
//   trsv_diag_kernel     one CTA, 128x128 triangular block staged in shared memory, one thread per
//                        unknown, one barrier per column
//   trsv_update_*_kernel rank-128 update of the remaining right-hand side, coalesced along rows of L
#include "nes_internal.h"
#include "ptx_util.cuh"

namespace nes {

constexpr int SV_NB = 128;
constexpr int SV_P = 130;  // smem pitch: 16B-aligned columns for bulk copies, <= 2-way conflicts on L'
constexpr int SV_SMEM = (SV_NB * SV_P + SV_NB) * 8 + 16;

// One CTA of 128 threads (thread = unknown).  The lower triangle of the 128x128 block arrives by
// one cp.async.bulk per column; the solve proceeds in four 32x32 sub-blocks, each by ONE WARP with
// shuffles (lane = row), followed by a rank-32 update of the remaining unknowns by all warps.
__global__ void __launch_bounds__(SV_NB)
trsv_diag_kernel(const double* __restrict__ M, long long ld, int j0, int jb,
                 const double* __restrict__ dinv, double* __restrict__ x, int transposed, int brows) {
    extern __shared__ __align__(16) double S[];
    M += blockIdx.y * brows;  // batched mode: blockIdx.y = problem, rows offset by brows
    dinv += blockIdx.y * brows;
    x += blockIdx.y * brows;
    double* xs = S + SV_NB * SV_P;
    uint64_t* bar = reinterpret_cast<uint64_t*>(xs + SV_NB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jbp = (jb + 1) & ~1;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        uint32_t total = 0;
        for (int cc = 0; cc < jb; ++cc) total += (uint32_t)(jbp - (cc & ~1)) * 8u;
        mbar_expect_tx(bar, total);
    }
    __syncthreads();
    if (tid < jb) {
        const int r0 = tid & ~1;
        bulk_load_1d(S + r0 + tid * SV_P, M + (j0 + r0) + (long long)(j0 + tid) * ld,
                     (uint32_t)(jbp - r0) * 8u, bar);
    }
    double v = 0.0, di = 1.0;
    if (tid < jb) {
        v = x[j0 + tid];
        di = dinv[j0 + tid];
    }
    mbar_wait(bar, 0);
    if (!transposed) {
        for (int sb = 0; sb < SV_NB / 32; ++sb) {
            const int base = 32 * sb;
            if (base >= jb) break;
            if (warp == sb) {
#pragma unroll 8
                for (int cc = 0; cc < 32; ++cc) {
                    const double xc = __shfl_sync(0xffffffffu, v * di, cc);
                    if (lane == cc) v = xc;
                    else if (lane > cc) v = fma(-S[(base + lane) + (base + cc) * SV_P], xc, v);
                }
                xs[base + lane] = v;
            }
            __syncthreads();
            if (tid >= base + 32) {
                double acc = 0.0;
#pragma unroll 8
                for (int cc = 0; cc < 32; ++cc) acc = fma(S[tid + (base + cc) * SV_P], xs[base + cc], acc);
                v -= acc;
            }
        }
    } else {
        for (int sb = SV_NB / 32 - 1; sb >= 0; --sb) {
            const int base = 32 * sb;
            if (base >= jb) continue;
            if (warp == sb) {
#pragma unroll 8
                for (int cc = 31; cc >= 0; --cc) {
                    const double xc = __shfl_sync(0xffffffffu, v * di, cc);
                    if (lane == cc) v = xc;
                    else if (lane < cc && base + cc < jb)
                        v = fma(-S[(base + cc) + (base + lane) * SV_P], xc, v);
                }
                xs[base + lane] = v;
            }
            __syncthreads();
            if (tid < base) {
                double acc = 0.0;
#pragma unroll 8
                for (int cc = 0; cc < 32; ++cc)
                    if (base + cc < jb) acc = fma(S[(base + cc) + tid * SV_P], xs[base + cc], acc);
                v -= acc;
            }
        }
    }
    if (tid < jb) x[j0 + tid] = v;
}

// forward: x[r] -= sum_c L[r, j0+c] x[j0+c] for r >= r0.  CTA = 32 rows x 8 column groups.
__global__ void __launch_bounds__(256)
trsv_update_fwd_kernel(const double* __restrict__ M, long long ld, int j0, int jb, int r0, int m,
                       double* __restrict__ x, int brows) {
    __shared__ double part[8][33];
    __shared__ double xk[SV_NB];
    M += blockIdx.y * brows;
    x += blockIdx.y * brows;
    const int tid = threadIdx.x;
    const int lr = tid & 31, grp = tid >> 5;
    if (tid < jb) xk[tid] = x[j0 + tid];
    __syncthreads();
    const int r = r0 + blockIdx.x * 32 + lr;
    double acc = 0.0;
    if (r < m) {
        const double* Lr = M + r + (long long)j0 * ld;
        const int c_lo = grp * 16, c_hi = min(jb, c_lo + 16);
#pragma unroll 16
        for (int cc = c_lo; cc < c_hi; ++cc) acc = fma(Lr[(long long)cc * ld], xk[cc], acc);
    }
    part[grp][lr] = acc;
    __syncthreads();
    if (grp == 0 && r < m) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += part[q][lr];
        x[r] -= s;
    }
}

// backward: x[c] -= sum_r L[j0+r, c] x[j0+r] for c < j0.  One warp per column c.
__global__ void __launch_bounds__(256)
trsv_update_bwd_kernel(const double* __restrict__ M, long long ld, int j0, int jb,
                       double* __restrict__ x, int brows) {
    __shared__ double xk[SV_NB];
    M += blockIdx.y * brows;
    x += blockIdx.y * brows;
    const int tid = threadIdx.x;
    if (tid < jb) xk[tid] = x[j0 + tid];
    __syncthreads();
    const int lane = tid & 31;
    const int cc = blockIdx.x * 8 + (tid >> 5);
    if (cc >= j0) return;
    const double* Lc = M + j0 + (long long)cc * ld;
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < SV_NB / 32; ++q) {
        const int r = lane + 32 * q;
        if (r < jb) acc = fma(Lc[r], xk[r], acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) x[cc] -= acc;
}

// ---- dataflow TRSV: one cooperative launch per sweep ---------------------------------------------
// CTA c owns block rows c, c+G, ... (128 rows each).  For block row i it streams the blocks L(i,k)
// (forward; L(k,i)' backward) in dependency order, spinning on a per-block-row "ready" flag until the
// CTA that owns block k has published its piece of the solution, and finally solves the diagonal
// block and publishes its own piece.  The critical path per block row is one 128x128 block product
// plus the diagonal solve plus one flag hand-off instead of two kernel launches.  Flags carry an
// epoch number, so they never need resetting.  Co-residency of all CTAs is guaranteed by
// cudaLaunchCooperativeKernel (grid <= SM count).
constexpr int DF_Q = 4;                       // each block row is worked on by 4 x 128 threads
constexpr int DF_W = SV_NB / DF_Q;            // columns (rows) of a block handled per thread: 32
constexpr int DF_THREADS = DF_Q * SV_NB;      // 512
constexpr int DF_SMEM = (2 + DF_Q) * SV_NB * 8;

__device__ __forceinline__ void df_wait(volatile int* flag, int epoch) {
    if (threadIdx.x == 0) {
        while (*flag != epoch) {
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(DF_THREADS, 1)
trsv_dataflow_kernel(const double* __restrict__ M, long long ld, int m, const double* __restrict__ Winv,
                     double* __restrict__ x, int* __restrict__ flags, int epoch, int transposed) {
    extern __shared__ __align__(16) double S[];
    double* xs = S;                      // right-hand side of the diagonal block
    double* yk = xs + SV_NB;             // the published piece of block k
    double* part = yk + SV_NB;           // part[q][t]: partial sums of the thread quarters
    const int tid = threadIdx.x, t = tid & 127, qd = tid >> 7;
    const int nblk = (m + SV_NB - 1) / SV_NB;
    auto combine = [&](double v) {       // sum over the four quarters, valid in quarter 0
        part[qd * SV_NB + t] = v;
        __syncthreads();
        double s = 0.0;
        if (qd == 0) s = (part[t] + part[SV_NB + t]) + (part[2 * SV_NB + t] + part[3 * SV_NB + t]);
        return s;
    };
    for (int q = blockIdx.x; q < nblk; q += gridDim.x) {
        const int i = transposed ? nblk - 1 - q : q;   // backward sweep walks the block rows downwards
        const int j0 = i * SV_NB;
        const int jb = min(SV_NB, m - j0);
        // this thread's quarter of row t (forward) / column t (backward) of W = L_ii^-1, fetched while the
        // dependencies are still in flight
        double wreg[DF_W];
        {
            const double* Wb = Winv + (size_t)i * SV_NB * SV_NB;
            if (!transposed) {
#pragma unroll
                for (int cc = 0; cc < DF_W; ++cc) wreg[cc] = Wb[t + (qd * DF_W + cc) * SV_NB];
            } else {
#pragma unroll
                for (int r = 0; r < DF_W; ++r) wreg[r] = Wb[(qd * DF_W + r) + t * SV_NB];
            }
        }
        __syncthreads();
        double acc = 0.0;
        // Each thread streams its quarter (32 values) of every 128x128 block.  The values of the NEXT
        // block are fetched into registers before spinning on that block's flag, so the DRAM/L2 latency
        // of L overlaps the wait and only 32 FMAs remain on the critical path after a flag flips.
        double lreg[DF_W];
        if (!transposed) {
            const double* Lr = M + j0 + t;  // row of this thread
            auto preload = [&](int k) {
                if (t < jb) {
                    const double* Lb = Lr + (long long)(k * SV_NB + qd * DF_W) * ld;
#pragma unroll
                    for (int cc = 0; cc < DF_W; ++cc) lreg[cc] = Lb[(long long)cc * ld];
                }
            };
            if (i > 0) preload(0);
            for (int k = 0; k < i; ++k) {
                df_wait(flags + k, epoch);
                if (tid < SV_NB) yk[tid] = __ldcg(x + k * SV_NB + tid);
                __syncthreads();
                if (t < jb) {
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int cc = 0; cc < DF_W; cc += 2) {
                        a0 = fma(lreg[cc], yk[qd * DF_W + cc], a0);
                        a1 = fma(lreg[cc + 1], yk[qd * DF_W + cc + 1], a1);
                    }
                    acc += a0 + a1;
                }
                if (k + 1 < i) preload(k + 1);
                __syncthreads();
            }
        } else {
            // Backward: the block L(k,i) is read ALONG ITS COLUMNS (rows contiguous): warp w owns columns
            // 8w .. 8w+7 of block i, lane l the rows l, l+32, l+64, l+96 of block k, so every load
            // instruction covers 256 contiguous bytes.  (One thread per column, 32 rows each, made every
            // load touch 32 different lines: 657 us per sweep under ncu against 262 us forward.)  The
            // lane-private partial sums are combined once per block row.
            const int warp = tid >> 5, lane = tid & 31;
            double acc8[8];
#pragma unroll
            for (int cI = 0; cI < 8; ++cI) acc8[cI] = 0.0;
            auto preload = [&](int k) {
                const int row0 = k * SV_NB + lane;
#pragma unroll
                for (int cI = 0; cI < 8; ++cI) {
                    const int col = j0 + 8 * warp + cI;
                    const double* Lc = M + (long long)min(col, m - 1) * ld;
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const int row = row0 + 32 * h;
                        lreg[cI * 4 + h] = (col < m && row < m) ? Lc[row] : 0.0;
                    }
                }
            };
            if (i < nblk - 1) preload(nblk - 1);
            for (int k = nblk - 1; k > i; --k) {
                df_wait(flags + k, epoch);
                const int kb = min(SV_NB, m - k * SV_NB);
                if (tid < SV_NB) yk[tid] = (tid < kb) ? __ldcg(x + k * SV_NB + tid) : 0.0;
                __syncthreads();
#pragma unroll
                for (int cI = 0; cI < 8; ++cI) {
                    double a0 = fma(lreg[cI * 4 + 0], yk[lane], 0.0);
                    double a1 = fma(lreg[cI * 4 + 1], yk[lane + 32], 0.0);
                    a0 = fma(lreg[cI * 4 + 2], yk[lane + 64], a0);
                    a1 = fma(lreg[cI * 4 + 3], yk[lane + 96], a1);
                    acc8[cI] += a0 + a1;
                }
                if (k - 1 > i) preload(k - 1);
                __syncthreads();
            }
            // sum over the lanes (fixed xor tree), lane cI keeps column 8 warp + cI
            double mine = 0.0;
#pragma unroll
            for (int cI = 0; cI < 8; ++cI) {
                double v8 = acc8[cI];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(0xffffffffu, v8, o);
                if (lane == cI) mine = v8;
            }
            __syncthreads();  // part[] may still be read by the previous block row's combine
            if (lane < 8) part[8 * warp + lane] = mine;
            __syncthreads();
            acc = (qd == 0) ? part[t] : 0.0;   // combine() below then returns exactly this total
            __syncthreads();
        }
        // right-hand side of the diagonal block, then y_i = W r (forward) or W' r (backward): a triangular
        // matvec, 32 FMAs per thread, instead of a 128-step substitution on the critical path
        const double tot = combine(acc);
        if (qd == 0) xs[t] = (t < jb) ? x[j0 + t] - tot : 0.0;
        __syncthreads();
        double v;
        {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int cc = 0; cc < DF_W; cc += 2) {
                a0 = fma(wreg[cc], xs[qd * DF_W + cc], a0);
                a1 = fma(wreg[cc + 1], xs[qd * DF_W + cc + 1], a1);
            }
            v = a0 + a1;
        }
        v = combine(v);
        if (qd == 0 && t < jb) x[j0 + t] = v;
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile int*>(flags + i) = epoch;
    }
}

static int dense_trsv_dataflow(nes_ctx* c, const double* M, long long ld, int m, const double* Winv,
                               double* d_x, int* d_flags, int* epoch) {
    static PerDeviceOnce once;
    int dev;
    if (once.begin(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(trsv_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             DF_SMEM);
        once.finish(dev, e == cudaSuccess);
        NES_CUDA(c, e);
    }
    const int nblk = (m + SV_NB - 1) / SV_NB;
    const int grid = nblk < c->num_sms ? nblk : c->num_sms;
    for (int transposed = 0; transposed < 2; ++transposed) {
        int ep = ++(*epoch);
        void* args[] = {(void*)&M, (void*)&ld, (void*)&m, (void*)&Winv, (void*)&d_x, (void*)&d_flags,
                        (void*)&ep, (void*)&transposed};
        NES_CUDA(c, cudaLaunchCooperativeKernel((const void*)trsv_dataflow_kernel, dim3(grid), dim3(DF_THREADS),
                                                args, DF_SMEM, c->stream));
        ++c->launches;
    }
    return 0;
}

// Both sweeps for `nbatch` stacked problems (brows = row stride between problems, also the stride of
// the right-hand sides and of dinv).  nbatch = 1, brows = 0 is the single-matrix case.
int dense_trsv_sweeps(nes_ctx* c, const double* M, long long ld, int m, const double* dinv, double* d_x,
                      int nbatch, int brows) {
    static PerDeviceOnce once;
    int dev;
    if (once.begin(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(trsv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SV_SMEM);
        once.finish(dev, e == cudaSuccess);
        NES_CUDA(c, e);
    }
    // L y = b
    for (int j0 = 0; j0 < m; j0 += SV_NB) {
        const int jb = (m - j0 < SV_NB) ? m - j0 : SV_NB;
        trsv_diag_kernel<<<dim3(1, nbatch), SV_NB, SV_SMEM, c->stream>>>(M, ld, j0, jb, dinv, d_x, 0, brows);
        NES_CHECK_LAUNCH(c);
        const int r0 = j0 + jb;
        if (r0 < m) {
            trsv_update_fwd_kernel<<<dim3((m - r0 + 31) / 32, nbatch), 256, 0, c->stream>>>(M, ld, j0, jb, r0,
                                                                                           m, d_x, brows);
            NES_CHECK_LAUNCH(c);
        }
    }
    // L' x = y
    const int nblk = (m + SV_NB - 1) / SV_NB;
    for (int k = nblk - 1; k >= 0; --k) {
        const int j0 = k * SV_NB;
        const int jb = (m - j0 < SV_NB) ? m - j0 : SV_NB;
        trsv_diag_kernel<<<dim3(1, nbatch), SV_NB, SV_SMEM, c->stream>>>(M, ld, j0, jb, dinv, d_x, 1, brows);
        NES_CHECK_LAUNCH(c);
        if (j0 > 0) {
            trsv_update_bwd_kernel<<<dim3((j0 + 7) / 8, nbatch), 256, 0, c->stream>>>(M, ld, j0, jb, d_x,
                                                                                     brows);
            NES_CHECK_LAUNCH(c);
        }
    }
    return 0;
}

int dense_solve_inplace(nes_ctx* c, nes_factor* L, double* d_x) {
    StageTimer timer(c, NES_STAGE_SOLVE);
    if (L->d_flags && L->d_Winv)
        return dense_trsv_dataflow(c, L->d_M, (long long)L->ld, (int)L->m, L->d_Winv, d_x, L->d_flags,
                                   &L->flag_epoch);
    return dense_trsv_sweeps(c, L->d_M, (long long)L->ld, (int)L->m, L->d_dinv, d_x, 1, 0);
}

}  // namespace nes
