// In-shared-memory Cholesky of one diagonal block (<= 128 x 128, column-major pitch CH_P), shared by
// the dense panel kernel (dense_chol.cu) and the supernode kernel (sparse_chol.cu).
//   8-column sub-panels: every row-owning thread factors the 8x8 pivot block redundantly in registers
//   (the first version used one warp + shuffles on a 16x16 pivot: ~490 cycles per column measured;
//   the register version's chain is rsqrt + mul + fma) and solves its own row; the trailing rank-8
//   update runs on a 16x16 thread grid with interleaved 8x8 register tiles.  256 threads.  A non-positive (or NaN) pivot records
//   info = {NES_NOT_POSDEF, col_base + column} once and continues with a unit pivot.
#pragma once
#include "../../include/nes.h"
#include "ptx_util.cuh"

namespace nes {

#ifdef POTRF_PROFILE
__device__ long long potrf_prof[8];
#endif

constexpr int CH_NB = 128;
constexpr int CH_P = 128;  // smem pitch of the diagonal block (column-major, dense TMA box)

// PITCH = shared-memory pitch of the block (CH_P for the 128 x 128 kernels, 64 for the batched 64 x 64 blocks)
// NA = 16-row groups of the largest trailing block (8 for 128 x 128, 4 for 64 x 64: a quarter of the accumulators)
template <int PITCH = CH_P, int NA = 8>
__device__ __forceinline__ void potrf_block_smem(double* S, double* dinv, int jb, double dbound,
                                                 int* __restrict__ info, int col_base) {
    const int tid = threadIdx.x;
    constexpr int W = 8;  // sub-panel width
#ifdef POTRF_PROFILE
    long long t_piv = 0, t_upd = 0, t0 = clock64();
#endif
    for (int c0 = 0; c0 < jb; c0 += W) {
        const int w = min(W, jb - c0);
        // (1)+(2) every thread that owns a row >= c0 factors the w x w pivot block REDUNDANTLY in its own
        // registers (no shuffles, no barrier between the pivot and the row solves: the per-column chain
        // is rsqrt + mul + fma) and then solves its row against it.
        const int r = c0 + tid;
#ifdef POTRF_PROFILE
        long long q0 = clock64(), q1 = q0, q2 = q0, q3 = q0;
#endif
        if (r < jb) {
            double P[W][W];   // lower triangle of the pivot block, then of its factor
            double ri[W];
            double x[W];
            if (w == W) {  // full sub-panel: unconditional broadcast loads
#pragma unroll
                for (int cc = 0; cc < W; ++cc)
#pragma unroll
                    for (int rr = cc; rr < W; ++rr) P[rr][cc] = S[(c0 + rr) + (c0 + cc) * PITCH];
#pragma unroll
                for (int cc = 0; cc < W; ++cc) x[cc] = S[r + (c0 + cc) * PITCH];
            } else {
#pragma unroll
                for (int cc = 0; cc < W; ++cc)
#pragma unroll
                    for (int rr = cc; rr < W; ++rr)
                        P[rr][cc] = (rr < w && cc < w) ? S[(c0 + rr) + (c0 + cc) * PITCH] : (rr == cc ? 1.0 : 0.0);
#pragma unroll
                for (int cc = 0; cc < W; ++cc) x[cc] = (cc < w) ? S[r + (c0 + cc) * PITCH] : 0.0;
            }
#ifdef POTRF_PROFILE
            q1 = clock64();
#endif
#pragma unroll
            for (int cc = 0; cc < W; ++cc) {
                double d = P[cc][cc];
                // CHOLMOD dbound for LL': L_jj is not allowed below dbound (0 = off)
                if (dbound > 0.0 && d < dbound * dbound) d = dbound * dbound;
                if (!(d > 0.0)) {  // also catches NaN
                    if (tid == 0 && cc < w && info[0] == 0) {
                        info[0] = NES_NOT_POSDEF;
                        info[1] = col_base + c0 + cc;
                    }
                    d = 1.0;
                }
                // rsqrt is the latency floor of the pivot chain (66 cycles on B200); it is good to 1 ulp,
                // so L_jj = d * rsqrt(d) is within 2 ulp of sqrt(d)
                ri[cc] = rsqrt(d);
                P[cc][cc] = d * ri[cc];
#pragma unroll
                for (int rr = cc + 1; rr < W; ++rr) P[rr][cc] *= ri[cc];
#pragma unroll
                for (int c2 = cc + 1; c2 < W; ++c2)
#pragma unroll
                    for (int rr = c2; rr < W; ++rr) P[rr][c2] = fma(-P[rr][cc], P[c2][cc], P[rr][c2]);
            }
#ifdef POTRF_PROFILE
            q2 = clock64();
#endif
            // x L_p' = a.  For a row INSIDE the pivot block the same recurrence yields row tid of the
            // factor in x[0..tid] (entries past the diagonal are junk and not stored), so every thread
            // takes one path: no divergence in warp 0.
#pragma unroll
            for (int cc = 0; cc < W; ++cc) {
                double acc = x[cc];
#pragma unroll
                for (int p = 0; p < cc; ++p) acc = fma(-x[p], P[cc][p], acc);
                x[cc] = acc * ri[cc];
            }
            const int last = (tid < w) ? tid : w - 1;
#pragma unroll
            for (int cc = 0; cc < W; ++cc)
                if (cc <= last) S[r + (c0 + cc) * PITCH] = x[cc];
            if (tid == 0) {
#pragma unroll
                for (int cc = 0; cc < W; ++cc)
                    if (cc < w) dinv[c0 + cc] = ri[cc];
            }
        }
#ifdef POTRF_PROFILE
        q3 = clock64();
        if (tid == 0 && c0 == 0) { potrf_prof[2] = q1 - q0; potrf_prof[3] = q2 - q1; potrf_prof[4] = q3 - q2; }
        if (tid == 96 && c0 == 0) { potrf_prof[6] = q3 - q0; }
        if (tid == 255 && c0 == 0) { potrf_prof[7] = q3 - q0; }
#endif
        __syncthreads();
#ifdef POTRF_PROFILE
        { long long t1 = clock64(); t_piv += t1 - t0; t0 = t1; if (tid == 0 && c0 == 0) potrf_prof[5] = t_piv; }
#endif
        const int base = c0 + w;
        const int T = jb - base;
        if (T <= 0) break;
        // (3) trailing rank-w update of the lower triangle, 16x16 thread grid, interleaved 8x8 tiles.
        // Loads are unconditional: rows past T alias the top of the next column (finite, and the
        // products land in accumulators that are never stored).
        {
            const int ti = tid & 15, tj = tid >> 4;
            double acc[NA][NA];
#pragma unroll
            for (int a_ = 0; a_ < NA; ++a_)
#pragma unroll
                for (int b_ = 0; b_ < NA; ++b_) acc[a_][b_] = 0.0;
            const double* colbase = S + c0 * PITCH + base;
            const int na = (T + 15) >> 4;  // 16-row groups in the trailing block (uniform)
            if (w == W) {
#pragma unroll
                for (int p = 0; p < W; ++p) {
                    const double* col = colbase + p * PITCH;
                    double xi[NA], xj[NA];
#pragma unroll
                    for (int a_ = 0; a_ < NA; ++a_) {
                        if (a_ < na) {
                            xi[a_] = col[ti + 16 * a_];
                            xj[a_] = col[tj + 16 * a_];
                        }
                    }
#pragma unroll
                    for (int a_ = 0; a_ < NA; ++a_) {
                        if (a_ < na) {
#pragma unroll
                            for (int b_ = 0; b_ <= a_; ++b_)
                                acc[a_][b_] = fma(xi[a_], xj[b_], acc[a_][b_]);
                        }
                    }
                }
            } else {
                for (int p = 0; p < w; ++p) {
                    const double* col = colbase + p * PITCH;
#pragma unroll
                    for (int a_ = 0; a_ < NA; ++a_) {
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
            for (int a_ = 0; a_ < NA; ++a_) {
                if (a_ < na) {
#pragma unroll
                    for (int b_ = 0; b_ <= a_; ++b_) {
                        const int i = ti + 16 * a_, j = tj + 16 * b_;
                        if (i < T && j <= i) S[(base + i) + (base + j) * PITCH] -= acc[a_][b_];
                    }
                }
            }
        }
        __syncthreads();
#ifdef POTRF_PROFILE
        { long long t1 = clock64(); t_upd += t1 - t0; t0 = t1; }
#endif
    }
#ifdef POTRF_PROFILE
    if (tid == 0) { potrf_prof[0] = t_piv; potrf_prof[1] = t_upd; }
#endif
}


// X L' = B for a slab of up to 64 rows against a diagonal block of nc <= 128 columns, all operands in
// shared memory:  Ls[c + p*128] = L(c, p) (zero above the diagonal), Xs[p*64 + row], dv[c] = 1/L(c, c).
// 256 threads = 4 threads per row (q = tid >> 6): within every 32-column block each thread owns 8
// columns; contributions of the columns already solved are applied by all four in parallel, the 8x8
// triangular pieces are solved by their owner.  True substitution (no inverse): the factorization's
// backward-error bound is untouched.  ~3.3x fewer FMAs on the per-row critical path than one thread
// per row.
template <int LP, int XP, bool kMma>
__device__ __forceinline__ void trsm_slab_smem_t(const double* Ls, double* Xs, const double* dv, int nc,
                                                 int nrows) {
    const int tid = threadIdx.x, r = tid & 63, q = tid >> 6;
    const bool active = r < nrows;
    for (int cb = 0; cb < nc; cb += 32) {
        const int c0 = cb + 8 * q;  // my 8 columns of this block
        if (kMma && cb > 0) {
            // X[:, cb:cb+32] -= X[:, 0:cb] L[cb:cb+32, 0:cb]' on the FP64 tensor cores: warp w owns rows
            // 8w..8w+7 and the four 8-column tiles; with pitches = 4 (mod 16) the fragment loads are
            // conflict-free.  (The FMA version of this product needs 5 shared loads per 8 FMAs.)
            const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
            double acc[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = 0.0;
            const double* xa = Xs + t * XP + 8 * warp + g;
            const double* lb = Ls + (cb + g) + t * LP;
            for (int k0 = 0; k0 < cb; k0 += 4) {
                const double a = xa[k0 * XP];
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[j][0], acc[j][1], a, lb[8 * j + k0 * LP]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) Xs[(cb + 8 * j + 2 * t + e) * XP + 8 * warp + g] -= acc[j][e];
            __syncthreads();
        }
        double b8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) b8[i] = (active && c0 + i < nc) ? Xs[(c0 + i) * XP + r] : 0.0;
        if (!kMma && active) {
            for (int p = 0; p < cb; ++p) {
                const double xp = Xs[p * XP + r];
                const double2* l2 = reinterpret_cast<const double2*>(Ls + c0 + p * LP);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double2 lv = l2[i];
                    b8[2 * i] = fma(-xp, lv.x, b8[2 * i]);
                    b8[2 * i + 1] = fma(-xp, lv.y, b8[2 * i + 1]);
                }
            }
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            const int s0 = cb + 8 * qq;  // first column of the sub-block being solved
            if (q == qq && active) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double xv = b8[i] * dv[min(s0 + i, 127)];
                    b8[i] = xv;
#pragma unroll
                    for (int i2 = i + 1; i2 < 8; ++i2)
                        b8[i2] = fma(-xv, Ls[min(s0 + i2, 127) + min(s0 + i, 127) * LP], b8[i2]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (s0 + i < nc) Xs[(s0 + i) * XP + r] = b8[i];
            }
            __syncthreads();
            if (q > qq && active && s0 < nc) {
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) {
                    const int p = s0 + pp;
                    if (p < nc) {
                        const double xp = Xs[p * XP + r];
                        const double2* l2 = reinterpret_cast<const double2*>(Ls + c0 + p * LP);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double2 lv = l2[i];
                            b8[2 * i] = fma(-xp, lv.x, b8[2 * i]);
                            b8[2 * i + 1] = fma(-xp, lv.y, b8[2 * i + 1]);
                        }
                    }
                }
            }
        }
    }
}

// dense-path layout: L block pitch 128, slab pitch 64 (TMA boxes), FMA product
__device__ __forceinline__ void trsm_slab_smem(const double* Ls, double* Xs, const double* dv, int nc,
                                               int nrows) {
    trsm_slab_smem_t<128, 64, false>(Ls, Xs, dv, nc, nrows);
}

}  // namespace nes
