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

// Both sweeps for `nbatch` stacked problems (brows = row stride between problems, also the stride of
// the right-hand sides and of dinv).  nbatch = 1, brows = 0 is the single-matrix case.
int dense_trsv_sweeps(nes_ctx* c, const double* M, long long ld, int m, const double* dinv, double* d_x,
                      int nbatch, int brows) {
    static bool configured = false;
    if (!configured) {
        NES_CUDA(c, cudaFuncSetAttribute(trsv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SV_SMEM));
        configured = true;
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
    return dense_trsv_sweeps(c, L->d_M, (long long)L->ld, (int)L->m, L->d_dinv, d_x, 1, 0);
}

}  // namespace nes
