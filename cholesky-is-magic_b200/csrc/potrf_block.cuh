// In-shared-memory Cholesky of one diagonal block (<= 128 x 128, column-major pitch CH_P), shared by
// the dense panel kernel (dense_chol.cu) and the supernode kernel (sparse_chol.cu).
//   16-column sub-panels: the 16x16 pivot block is factored by ONE WARP with shuffles (lane = row),
//   the rows below by one thread per row, the trailing rank-16 update on a 16x16 thread grid with
//   interleaved 7x7 register tiles.  256 threads.  A non-positive (or NaN) pivot records
//   info = {NES_NOT_POSDEF, col_base + column} once and continues with a unit pivot.
#pragma once
#include "../../include/nes.h"

namespace nes {

constexpr int CH_NB = 128;
constexpr int CH_W = 16;
constexpr int CH_P = 128;  // smem pitch of the diagonal block (column-major, dense TMA box)

__device__ __forceinline__ void potrf_block_smem(double* S, double* dinv, int jb, double dbound,
                                                 int* __restrict__ info, int col_base) {
    const int tid = threadIdx.x;
    const int j0 = col_base;
    for (int c0 = 0; c0 < jb; c0 += CH_W) {
        const int w = min(CH_W, jb - c0);
        // (1) pivot block, one warp, lane = row
        if (tid < 32) {
            const int lane = tid;
            double a[CH_W];
#pragma unroll
            for (int cc = 0; cc < CH_W; ++cc)
                a[cc] = (lane < w && cc <= lane) ? S[(c0 + lane) + (c0 + cc) * CH_P] : 0.0;
#pragma unroll
            for (int cc = 0; cc < CH_W; ++cc) {
                if (cc < w) {
                    double d = __shfl_sync(0xffffffffu, a[cc], cc);
                    // CHOLMOD dbound for LL': L_jj is not allowed below dbound (0 = off)
                    if (dbound > 0.0 && d < dbound * dbound) d = dbound * dbound;
                    if (!(d > 0.0)) {  // also catches NaN
                        if (lane == 0 && info[0] == 0) {
                            info[0] = NES_NOT_POSDEF;
                            info[1] = j0 + c0 + cc;
                        }
                        d = 1.0;
                    }
                    // rsqrt is the latency floor of the pivot chain (66 cycles on B200); it is good to
                    // 1 ulp, so L_jj = d * rsqrt(d) is within 2 ulp of sqrt(d)
                    const double ri = rsqrt(d);
                    a[cc] = (lane == cc) ? d * ri : a[cc] * ri;
#pragma unroll
                    for (int c2 = cc + 1; c2 < CH_W; ++c2) {
                        const double l = __shfl_sync(0xffffffffu, a[cc], c2);
                        a[c2] = fma(-a[cc], l, a[c2]);
                    }
                    if (lane == 0) dinv[c0 + cc] = ri;
                }
            }
#pragma unroll
            for (int cc = 0; cc < CH_W; ++cc)
                if (lane < w && cc <= lane) S[(c0 + lane) + (c0 + cc) * CH_P] = a[cc];
        }
        __syncthreads();
        const int base = c0 + w;
        const int T = jb - base;
        if (T <= 0) break;
        // (2) rows below the pivot block: x L_d' = a, one thread per row
        if (tid < T) {
            const int r = base + tid;
            double x[CH_W];
#pragma unroll
            for (int cc = 0; cc < CH_W; ++cc) x[cc] = (cc < w) ? S[r + (c0 + cc) * CH_P] : 0.0;
#pragma unroll
            for (int cc = 0; cc < CH_W; ++cc) {
                if (cc < w) {
                    double acc = x[cc];
#pragma unroll
                    for (int p = 0; p < cc; ++p)
                        acc = fma(-x[p], S[(c0 + cc) + (c0 + p) * CH_P], acc);
                    x[cc] = acc * dinv[c0 + cc];
                }
            }
#pragma unroll
            for (int cc = 0; cc < CH_W; ++cc)
                if (cc < w) S[r + (c0 + cc) * CH_P] = x[cc];
        }
        __syncthreads();
        // (3) trailing rank-w update of the lower triangle, 16x16 thread grid, interleaved 7x7 tiles.
        // Loads are unconditional: rows past T alias the top of the next column (finite, and the
        // products land in accumulators that are never stored).
        {
            const int ti = tid & 15, tj = tid >> 4;
            double acc[7][7];
#pragma unroll
            for (int a_ = 0; a_ < 7; ++a_)
#pragma unroll
                for (int b_ = 0; b_ < 7; ++b_) acc[a_][b_] = 0.0;
            const double* colbase = S + c0 * CH_P + base;
            const int na = (T + 15) >> 4;  // 16-row groups in the trailing block (uniform)
            if (w == CH_W) {
#pragma unroll 4
                for (int p = 0; p < CH_W; ++p) {
                    const double* col = colbase + p * CH_P;
                    double xi[7], xj[7];
#pragma unroll
                    for (int a_ = 0; a_ < 7; ++a_) {
                        if (a_ < na) {
                            xi[a_] = col[ti + 16 * a_];
                            xj[a_] = col[tj + 16 * a_];
                        }
                    }
#pragma unroll
                    for (int a_ = 0; a_ < 7; ++a_) {
                        if (a_ < na) {
#pragma unroll
                            for (int b_ = 0; b_ <= a_; ++b_)
                                acc[a_][b_] = fma(xi[a_], xj[b_], acc[a_][b_]);
                        }
                    }
                }
            } else {
                for (int p = 0; p < w; ++p) {
                    const double* col = colbase + p * CH_P;
#pragma unroll
                    for (int a_ = 0; a_ < 7; ++a_) {
                        if (a_ < na) {
                            const double xa = col[ti + 16 * a_];
#pragma unroll
                            for (int b_ = 0; b_ <= a_; ++b_)
                                acc[a_][b_] = fma(xa, col[tj + 16 * b_], acc[a_][b_]);
                        }
                    }
                }
            }
#pragma unroll
            for (int a_ = 0; a_ < 7; ++a_) {
                if (a_ < na) {
#pragma unroll
                    for (int b_ = 0; b_ <= a_; ++b_) {
                        const int i = ti + 16 * a_, j = tj + 16 * b_;
                        if (i < T && j <= i) S[(base + i) + (base + j) * CH_P] -= acc[a_][b_];
                    }
                }
            }
        }
        __syncthreads();
    }

}

}  // namespace nes
