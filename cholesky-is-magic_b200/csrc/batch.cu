// Batched small dense problems (BASELINE config 5: 1024 independent LPs of m = 256 solved by affine
// scaling).  The reference would run `affine-scaling` (affine-scaling.lisp:265-297) once per LP; here
// all LPs advance together and every kernel covers the whole batch:
//   * the B constraint matrices are stacked along the rows of ONE column-major matrix (row stride
//     mp = ceil128(m) per problem), so the TMA tensor maps, the DMMA formation / update kernel
//     (dmma_nt, batched tile enumeration), the diagonal-block, TRSM and TRSV kernels are the same code
//     as the single-matrix path with blockIdx.y (or the tile index) selecting the problem;
//   * formation: B * 3 lower tiles (m = 256) with K = n in one launch; Cholesky: 4 launches for the
//     whole batch (potrf, trsm, update, potrf); solves: 6 launches;
//   * per-problem reductions (one CTA per LP) return scalars to the host, which runs the control flow
//     of one-iteration / one-affine-scaling-iteration (:165-263) per LP and sends back a mode and a
//     step per problem.
#include <cmath>
#include <vector>

#include <cstdlib>

#include "dmma_nt.cuh"
#include "potrf_block.cuh"
#include "ipm_kernels.cuh"
#include "nes_internal.h"

using namespace nes;

struct nes_batch {
    int B = 0, m = 0, n = 0, mp = 0, np = 0;
    size_t ld = 0;            // = B * mp
    double* d_A = nullptr;    // (B*mp) x n
    double* d_M = nullptr;    // (B*mp) x mp  (per problem: lower triangle of M, then L)
    double* d_dinv = nullptr; // B*mp
    int* d_info = nullptr;    // 2*B
    int* d_mode = nullptr;    // B
    double* d_step = nullptr; // B
    double* d_block = nullptr;
    double *c, *l, *u, *x, *slack, *sc, *g, *theta;  // B*np
    double *b, *r, *t;                               // B*mp
    double* d_scal = nullptr;                        // B*8
    CUtensorMap mapA, mapM, mapBlk, mapSlab;
};

enum { BM_SKIP = 0, BM_REPAIR = 1, BM_OPT = 2, BM_CENTER = 3 };

namespace nes {

// y[b] = alpha * A_b (s_b . x_b) + beta * y[b];  CTA = (256-row block, problem)
__global__ void __launch_bounds__(256)
batch_gemv_n_kernel(const double* __restrict__ A, size_t ld, int m, int n, int mp, int np,
                    const double* __restrict__ x, const double* __restrict__ s, double alpha, double beta,
                    double* __restrict__ y) {
    __shared__ double xs[512];
    const int bb = blockIdx.y;
    const int r = blockIdx.x * 256 + threadIdx.x;
    const double* Ab = A + (size_t)bb * mp;
    const double* xb = x + (size_t)bb * np;
    const double* sb = s ? s + (size_t)bb * np : nullptr;
    double acc = 0.0;
    for (int k0 = 0; k0 < n; k0 += 512) {
        const int kn = min(512, n - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kn; i += 256) xs[i] = sb ? xb[k0 + i] * sb[k0 + i] : xb[k0 + i];
        __syncthreads();
        if (r < m) {
            const double* Ap = Ab + r + (size_t)k0 * ld;
#pragma unroll 8
            for (int k = 0; k < kn; ++k) acc = fma(Ap[(size_t)k * ld], xs[k], acc);
        }
    }
    if (r < m) {
        double* yp = y + (size_t)bb * mp + r;
        *yp = (beta == 0.0) ? alpha * acc : fma(alpha, acc, beta * *yp);
    }
}

// out[b][k] = alpha * s_k * sum_r A_b[r,k] x_b[r] + beta * out[b][k];  warp per (column, problem)
__global__ void __launch_bounds__(256)
batch_gemv_t_kernel(const double* __restrict__ A, size_t ld, int m, int n, int mp, int np,
                    const double* __restrict__ x, const double* __restrict__ s, double alpha, double beta,
                    double* __restrict__ y) {
    const int bb = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= n) return;
    const double* Ac = A + (size_t)bb * mp + (size_t)k * ld;
    const double* xb = x + (size_t)bb * mp;
    double acc = 0.0;
    for (int r = lane; r < m; r += 32) acc = fma(Ac[r], xb[r], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        if (s) acc *= s[(size_t)bb * np + k];
        double* yp = y + (size_t)bb * np + k;
        *yp = (beta == 0.0) ? alpha * acc : fma(alpha, acc, beta * *yp);
    }
}

// r = b - A x was formed by the GEMV; scal[b] = { |r|^2, c.x }
__global__ void __launch_bounds__(256)
batch_residual_scalars_kernel(int m, int n, int mp, int np, const double* __restrict__ r,
                              const double* __restrict__ x, const double* __restrict__ cvec,
                              double* __restrict__ scal) {
    __shared__ double buf[32];
    const int bb = blockIdx.x;
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) {
        const double v = r[(size_t)bb * mp + i];
        s0 = fma(v, v, s0);
    }
    for (int i = threadIdx.x; i < n; i += 256) s1 = fma(x[(size_t)bb * np + i], cvec[(size_t)bb * np + i], s1);
    double v = block_reduce(s0, RED_SUM, buf);
    if (threadIdx.x == 0) scal[bb * 8 + 0] = v;
    v = block_reduce(s1, RED_SUM, buf);
    if (threadIdx.x == 0) scal[bb * 8 + 1] = v;
}

// per problem, by mode: slack (cap sqrt(1e8) for repair, 1e8 otherwise), theta = slack^2,
// sc = slack * (-(c | centering-direction)), scal[b][7] = min slack
__global__ void __launch_bounds__(256)
batch_slack_kernel(int n, int np, const int* __restrict__ mode, const double* __restrict__ x,
                   const double* __restrict__ lo, const double* __restrict__ hi,
                   const double* __restrict__ cvec, double* __restrict__ slack, double* __restrict__ theta,
                   double* __restrict__ sc, double* __restrict__ scal) {
    __shared__ double buf[32];
    const int bb = blockIdx.x;
    const int md = mode[bb];
    const double cap = (md == BM_REPAIR) ? sqrt(1e8) : 1e8;
    double mn = INFINITY;
    for (int i = threadIdx.x; i < np; i += 256) {
        const size_t k = (size_t)bb * np + i;
        if (i >= n) {
            theta[k] = 0.0;
            continue;
        }
        const double xi = x[k], lb = lo[k], ub = hi[k];
        const double sl = fmin(cap, fmin(xi - lb, ub - xi));
        slack[k] = sl;
        theta[k] = sl * sl;
        mn = fmin(mn, sl);
        double d = cvec[k];
        if (md == BM_CENTER) {
            if (isinf(lb) && isinf(ub)) d = 0.0;
            else if ((xi - lb) < (ub - xi)) d = fmin(1.0, ub - xi);
            else d = fmax(-1.0, lb - xi);
        }
        sc[k] = sl * (-1.0 * d);
    }
    mn = block_reduce(mn, RED_MIN, buf);
    if (threadIdx.x == 0) scal[bb * 8 + 7] = mn;
}

// rhs of the normal equations: repair -> the residual r, otherwise AD sc (already in t)
__global__ void batch_select_rhs_kernel(int m, int mp, const int* __restrict__ mode,
                                        const double* __restrict__ r, double* __restrict__ t) {
    const int bb = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m && mode[bb] == BM_REPAIR) t[(size_t)bb * mp + i] = r[(size_t)bb * mp + i];
}

// dg: repair -> AD' t (in w), otherwise sc - AD' t;  g = dg * slack;
// scal[b] = { max-step, sum g^2, sum dg^2, sum g.c }
__global__ void __launch_bounds__(256)
batch_direction_kernel(int n, int np, const int* __restrict__ mode, const double* __restrict__ w,
                       const double* __restrict__ sc, const double* __restrict__ slack,
                       const double* __restrict__ x, const double* __restrict__ lo,
                       const double* __restrict__ hi, const double* __restrict__ cvec,
                       double* __restrict__ g, double* __restrict__ scal) {
    __shared__ double buf[32];
    const int bb = blockIdx.x;
    const int md = mode[bb];
    double mn = INFINITY, sg = 0.0, sd = 0.0, sgc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const size_t k = (size_t)bb * np + i;
        const double di = (md == BM_REPAIR) ? w[k] : sc[k] - w[k];
        const double gi = di * slack[k];
        g[k] = gi;
        sg = fma(gi, gi, sg);
        sd = fma(di, di, sd);
        sgc = fma(gi, cvec[k], sgc);
        if (gi < 0.0) mn = fmin(mn, (lo[k] - x[k]) / gi);
        else if (gi > 0.0) mn = fmin(mn, (hi[k] - x[k]) / gi);
    }
    double v = block_reduce(mn, RED_MIN, buf);
    if (threadIdx.x == 0) scal[bb * 8 + 2] = v;
    v = block_reduce(sg, RED_SUM, buf);
    if (threadIdx.x == 0) scal[bb * 8 + 3] = v;
    v = block_reduce(sd, RED_SUM, buf);
    if (threadIdx.x == 0) scal[bb * 8 + 4] = v;
    v = block_reduce(sgc, RED_SUM, buf);
    if (threadIdx.x == 0) scal[bb * 8 + 5] = v;
}

__global__ void batch_axpy_kernel(int n, int np, const double* __restrict__ step,
                                  const double* __restrict__ g, double* __restrict__ x) {
    const int bb = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double st = step[bb];
    if (i < n && st != 0.0) {
        const size_t k = (size_t)bb * np + i;
        x[k] = fma(st, g[k], x[k]);
    }
}

__global__ void batch_pack_rows_kernel(const double* __restrict__ src, size_t src_ld, int m, int n, int mp,
                                       int B, double* __restrict__ dst, size_t ld) {
    // src: B matrices m x n column-major, contiguous;  dst: stacked, problem b at rows [b*mp, b*mp+m)
    const size_t total = (size_t)B * m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t bb = idx / ((size_t)m * n);
        const size_t rem = idx - bb * (size_t)m * n;
        const size_t k = rem / m, r = rem - k * m;
        dst[bb * mp + r + k * ld] = src[idx];
    }
}

// form + factor all problems:  M_b = A_b diag(theta_b) A_b',  M_b = L_b L_b'
// ---- batched Cholesky on 64 x 64 diagonal blocks (round 2) -------------------------------------------
// The 128 x 128 panel kernels of the dense path need 133 KB (potrf) / 196 KB (trsm) of shared memory: one CTA
// per SM, and both are latency chains, so 1024 problems take 7 + 14 waves of 43 / 20 us (1.05 ms per batched
// factorization, 0.15 of the tensor roofline).  With 64 x 64 blocks a CTA needs 33 KB / 66 KB and two / three are
// resident per SM; the trailing updates stay on the DMMA kernel (K = 64).  MEASURED (1024 x 256 x 512, B200):
// 1.17 ms against 1.05 ms -- the potrf/trsm chains do get shorter, but the ten launches' trailing updates on
// 128 x 128 tiles with K = 64 pay the tile's fixed cost (operand fill + 256 KB read-modify-write) 5120 times.
// Kept opt-in (NES_BATCH_PANEL64=1) with its parity tests; what the batched factor needs is a fused
// trsm + update kernel on 64 x 64 output tiles, not smaller panels.
constexpr int B64 = 64;

__global__ void __launch_bounds__(256, 2)
potrf64_batch_kernel(double* __restrict__ M, long long ld, int i0, int jb, double* __restrict__ dinv_g,
                     double dbound, int* __restrict__ info, int brows) {
    __shared__ double S[B64 * B64];
    __shared__ double dv[B64];
    const int tid = threadIdx.x;
    const long long brow = (long long)blockIdx.x * brows;
    double* blk = M + brow + i0 + (long long)i0 * ld;
    for (int idx = tid; idx < B64 * B64; idx += 256) {
        const int cc = idx >> 6, r = idx & 63;
        S[idx] = (r < jb && cc < jb) ? blk[r + (long long)cc * ld] : (r == cc ? 1.0 : 0.0);
    }
    __syncthreads();
    potrf_block_smem<B64, 4>(S, dv, jb, dbound, info + 2 * blockIdx.x, i0);
    __syncthreads();
    for (int idx = tid; idx < B64 * B64; idx += 256) {
        const int cc = idx >> 6, r = idx & 63;
        if (r < jb && cc <= r) blk[r + (long long)cc * ld] = S[idx];
    }
    if (tid < jb) dinv_g[brow + i0 + tid] = dv[tid];
}

__global__ void __launch_bounds__(256)
trsm64_batch_kernel(double* __restrict__ M, long long ld, int i0, int jb, int m, const double* __restrict__ dinv_g,
                    int brows) {
    extern __shared__ double sm64[];
    double* Ls = sm64;                 // Ls[c + p*64] = L(c, p), zero above the diagonal
    double* Xs = Ls + B64 * B64;       // Xs[p*64 + row]
    double* dv = Xs + B64 * B64;
    const int tid = threadIdx.x;
    const long long brow = (long long)blockIdx.y * brows;
    const int row0 = i0 + jb + blockIdx.x * 64;
    const int nrows = min(64, m - row0);
    const double* Lg = M + brow + i0 + (long long)i0 * ld;
    double* Xg = M + brow + row0 + (long long)i0 * ld;
    for (int idx = tid; idx < B64 * B64; idx += 256) {
        const int p = idx >> 6, r = idx & 63;
        Ls[idx] = (r < jb && p < jb && r >= p) ? Lg[r + (long long)p * ld] : 0.0;
        Xs[idx] = (r < nrows && p < jb) ? Xg[r + (long long)p * ld] : 0.0;
    }
    if (tid < B64) dv[tid] = (tid < jb) ? dinv_g[brow + i0 + tid] : 1.0;
    __syncthreads();
    trsm_slab_smem_t<B64, B64, false>(Ls, Xs, dv, jb, nrows);
    __syncthreads();
    for (int idx = tid; idx < B64 * B64; idx += 256) {
        const int p = idx >> 6, r = idx & 63;
        if (r < nrows && p < jb) Xg[r + (long long)p * ld] = Xs[idx];
    }
}

static int batch_factor(nes_ctx* c, nes_batch* bt) {
    const int B = bt->B, m = bt->m, mp = bt->mp;
    const int tm = (m + NT_BM - 1) / NT_BM;
    {
        StageTimer t(c, NES_STAGE_FORM);
        NtArgs a{};
        a.C = bt->d_M;
        a.ldc = (long long)bt->ld;
        a.M = a.N = m;
        a.k0 = 0;
        a.K = bt->n;
        a.scale = bt->theta;
        a.alpha = 1.0;
        a.beta = 0.0;
        a.lower = 1;
        a.same_operand = 1;
        a.batch_tiles = tm * (tm + 1) / 2;
        a.batch_rows = mp;
        a.batch_scale_stride = bt->np;
        a.ntiles = B * a.batch_tiles;
        cudaError_t e = nt_launch(bt->mapA, bt->mapA, a, c->num_sms, c->stream);
        ++c->launches;
        if (e != cudaSuccess) return fail(c, NES_ERR_CUDA, "batched formation failed: %s", cudaGetErrorString(e));
    }
    StageTimer t(c, NES_STAGE_FACTOR);
    NES_CUDA(c, cudaMemsetAsync(bt->d_info, 0, 2 * B * sizeof(int), c->stream));
    // 128-column panels by default; NES_BATCH_PANEL64=1 selects the 64 x 64 block kernels above (measured: not faster)
    const bool wide = getenv("NES_BATCH_PANEL64") == nullptr;
    const int nb = wide ? 128 : B64;
    for (int i0 = 0; i0 < m; i0 += nb) {
        const int ib = (m - i0 < nb) ? m - i0 : nb;
        if (wide) {
            NES_TRY(chol_panel_launch(c, bt->mapBlk, bt->mapSlab, i0, ib, m, bt->d_dinv, bt->d_info, B, mp));
        } else {
            potrf64_batch_kernel<<<B, 256, 0, c->stream>>>(bt->d_M, (long long)bt->ld, i0, ib, bt->d_dinv, c->dbound,
                                                          bt->d_info, mp);
            NES_CHECK_LAUNCH(c);
            if (m - i0 - ib > 0) {
                constexpr int kTrsmSmem = (2 * B64 * B64 + B64) * 8;
                static PerDeviceOnce once;
                int dev;
                if (once.begin(&dev)) {
                    cudaError_t e = cudaFuncSetAttribute(trsm64_batch_kernel,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmem);
                    once.finish(dev, e == cudaSuccess);
                    NES_CUDA(c, e);
                }
                trsm64_batch_kernel<<<dim3((m - i0 - ib + 63) / 64, B), 256, kTrsmSmem, c->stream>>>(
                    bt->d_M, (long long)bt->ld, i0, ib, m, bt->d_dinv, mp);
                NES_CHECK_LAUNCH(c);
            }
        }
        const int rest = m - i0 - ib;
        if (rest > 0) {
            const int tr = (rest + NT_BM - 1) / NT_BM;
            NtArgs a{};
            a.C = bt->d_M + (i0 + ib) + (long long)(i0 + ib) * (long long)bt->ld;
            a.ldc = (long long)bt->ld;
            a.M = a.N = rest;
            a.rowA0 = a.rowB0 = i0 + ib;
            a.k0 = i0;
            a.K = ib;
            a.alpha = -1.0;
            a.beta = 1.0;
            a.lower = 1;
            a.same_operand = 1;
            a.batch_tiles = tr * (tr + 1) / 2;
            a.batch_rows = mp;
            a.ntiles = B * a.batch_tiles;
            cudaError_t e = nt_launch(bt->mapM, bt->mapM, a, c->num_sms, c->stream);
            ++c->launches;
            if (e != cudaSuccess) return fail(c, NES_ERR_CUDA, "batched update failed: %s", cudaGetErrorString(e));
        }
    }
    return 0;
}

static int batch_solve(nes_ctx* c, nes_batch* bt, double* d_rhs) {
    StageTimer t(c, NES_STAGE_SOLVE);
    return dense_trsv_sweeps(c, bt->d_M, (long long)bt->ld, bt->m, bt->d_dinv, d_rhs, bt->B, bt->mp);
}

}  // namespace nes

extern "C" {

nes_batch* nes_batch_create(const double* A_all, int B, int m, int n, const double* c_all,
                            const double* b_all, const double* l_all, const double* u_all,
                            const double* x_all, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (!A_all || B <= 0 || m <= 0 || n <= 0) {
        fail(c, NES_ERR_INVALID, "nes_batch_create: bad arguments");
        return nullptr;
    }
    nes_batch* bt = new nes_batch();
    bt->B = B; bt->m = m; bt->n = n;
    bt->mp = (m + 127) / 128 * 128;
    bt->np = (n + 15) / 16 * 16;
    bt->ld = (size_t)B * bt->mp;
    const size_t nA = bt->ld * n, nM = bt->ld * bt->mp;
    const size_t nv = (size_t)B * bt->np, mv = (size_t)B * bt->mp;
    bt->d_A = static_cast<double*>(dev_alloc(c, nA * sizeof(double)));
    bt->d_M = static_cast<double*>(dev_alloc(c, nM * sizeof(double)));
    bt->d_dinv = static_cast<double*>(dev_alloc(c, (mv + 16) * sizeof(double)));
    bt->d_info = static_cast<int*>(dev_alloc(c, 2 * B * sizeof(int)));
    bt->d_mode = static_cast<int*>(dev_alloc(c, B * sizeof(int)));
    bt->d_step = static_cast<double*>(dev_alloc(c, B * sizeof(double)));
    bt->d_scal = static_cast<double*>(dev_alloc(c, (size_t)B * 8 * sizeof(double)));
    bt->d_block = static_cast<double*>(dev_alloc(c, (8 * nv + 3 * mv + 64) * sizeof(double)));
    double* stage = static_cast<double*>(dev_alloc(c, (size_t)B * m * n * sizeof(double)));
    if (!bt->d_A || !bt->d_M || !bt->d_dinv || !bt->d_info || !bt->d_mode || !bt->d_step || !bt->d_scal ||
        !bt->d_block || !stage) {
        dev_free(c, stage);
        nes_batch_free(&bt, c);
        return nullptr;
    }
    double* p = bt->d_block;
    auto take = [&](size_t len) { double* q = p; p += len; return q; };
    bt->c = take(nv); bt->l = take(nv); bt->u = take(nv); bt->x = take(nv); bt->slack = take(nv);
    bt->sc = take(nv); bt->g = take(nv); bt->theta = take(nv);
    bt->b = take(mv); bt->r = take(mv); bt->t = take(mv);
    bool ok = cudaMemsetAsync(bt->d_A, 0, nA * sizeof(double), c->stream) == cudaSuccess &&
              cudaMemsetAsync(bt->d_M, 0, nM * sizeof(double), c->stream) == cudaSuccess &&
              cudaMemsetAsync(bt->d_block, 0, (8 * nv + 3 * mv + 64) * sizeof(double), c->stream) == cudaSuccess &&
              upload(c, stage, A_all, (size_t)B * m * n * sizeof(double)) == 0;
    if (ok) {
        batch_pack_rows_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(stage, m, m, n, bt->mp, B, bt->d_A, bt->ld);
        ++c->launches;
        ok = cudaGetLastError() == cudaSuccess;
    }
    // vectors: per-problem slices of length n (m) at stride np (mp)
    auto up2 = [&](double* dst, const double* src, int len, int stride) {
        if (!src) return true;
        return cudaMemcpy2DAsync(dst, stride * sizeof(double), src, len * sizeof(double), len * sizeof(double), B,
                                 cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
    };
    ok = ok && up2(bt->c, c_all, n, bt->np) && up2(bt->l, l_all, n, bt->np) && up2(bt->u, u_all, n, bt->np) &&
         up2(bt->x, x_all, n, bt->np) && up2(bt->b, b_all, m, bt->mp) &&
         cudaStreamSynchronize(c->stream) == cudaSuccess;
    dev_free(c, stage);
    ok = ok && make_operand_map(&bt->mapA, bt->d_A, (long long)bt->ld, n, (long long)bt->ld) == 0 &&
         make_operand_map(&bt->mapM, bt->d_M, (long long)bt->ld, bt->mp, (long long)bt->ld) == 0 &&
         make_operand_map(&bt->mapBlk, bt->d_M, (long long)bt->ld, bt->mp, (long long)bt->ld, 128, 128) == 0 &&
         make_operand_map(&bt->mapSlab, bt->d_M, (long long)bt->ld, bt->mp, (long long)bt->ld, 64, 128) == 0;
    if (!ok) {
        fail(c, NES_ERR_CUDA, "nes_batch_create: setup failed");
        nes_batch_free(&bt, c);
        return nullptr;
    }
    return bt;
}

int nes_batch_free(nes_batch** bt, nes_ctx* c) {
    if (!c) return 0;
    if (!bt || !*bt) return 1;
    if (c->started) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
    }
    nes_batch* b = *bt;
    dev_free(c, b->d_A); dev_free(c, b->d_M); dev_free(c, b->d_dinv); dev_free(c, b->d_info);
    dev_free(c, b->d_mode); dev_free(c, b->d_step); dev_free(c, b->d_scal); dev_free(c, b->d_block);
    delete b;
    *bt = nullptr;
    return 1;
}

// Batched solve-dense (sparse-cholesky.lisp:409-431): for every problem x_b with
// (A_b diag(s_b))(A_b diag(s_b))' x_b = rhs_b.  s_all may be NULL (unscaled).  status[b] = 0 / 1.
int nes_batch_normal_solve(nes_batch* bt, const double* s_all, const double* rhs_all, double* x_all,
                           int* status, nes_ctx* c) {
    NES_ENTER(c);
    if (!bt || !rhs_all || !x_all) return fail(c, NES_ERR_INVALID, "nes_batch_normal_solve: null argument");
    const int B = bt->B, m = bt->m, n = bt->n;
    std::vector<double> th((size_t)B * bt->np, 0.0);
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < n; ++k) {
            const double s = s_all ? s_all[(size_t)b * n + k] : 1.0;
            th[(size_t)b * bt->np + k] = s * s;
        }
    NES_TRY(upload(c, bt->theta, th.data(), th.size() * sizeof(double)));
    NES_CUDA(c, cudaMemcpy2DAsync(bt->t, bt->mp * sizeof(double), rhs_all, m * sizeof(double), m * sizeof(double),
                                  B, cudaMemcpyHostToDevice, c->stream));
    NES_TRY(batch_factor(c, bt));
    NES_TRY(batch_solve(c, bt, bt->t));
    NES_CUDA(c, cudaMemcpy2DAsync(x_all, m * sizeof(double), bt->t, bt->mp * sizeof(double), m * sizeof(double),
                                  B, cudaMemcpyDeviceToHost, c->stream));
    std::vector<int> info(2 * B);
    NES_TRY(download(c, info.data(), bt->d_info, info.size() * sizeof(int)));
    int any = 0;
    for (int b = 0; b < B; ++b) {
        if (status) status[b] = info[2 * b];
        any |= info[2 * b];
    }
    c->status = any ? NES_NOT_POSDEF : 0;
    return 0;
}

// Batched affine-scaling (affine-scaling.lisp:265-297): all LPs advance together; iters[b], obj[b]
// (c'x) and res[b] (|b - Ax|) are per problem.  Returns 0; per-problem Cholesky failures stop that LP
// (" singular ", :178-181) and are flagged with iters[b] < 0.
int nes_batch_affine_solve(nes_batch* bt, int max_iter, int* iters, double* obj, double* res, nes_ctx* c) {
    NES_ENTER(c);
    if (!bt) return fail(c, NES_ERR_INVALID, "nes_batch_affine_solve: null state");
    const int B = bt->B, m = bt->m, n = bt->n, mp = bt->mp, np = bt->np;
    std::vector<int> mode(B), done(B, 0), cont(B, 1), it_count(B, 0), info(2 * B);
    std::vector<double> step(B), scal((size_t)B * 8), objv(B, 0.0), resv(B, 0.0);
    const double gamma = 0.9, thr = 1e-6 * (double)m;
    const dim3 gm((m + 255) / 256, B), gn((n + 255) / 256, B), gt((n + 7) / 8, B);
    auto direction_pass = [&]() -> int {
        // slack / theta / sc by mode; rhs; factor; solve; dg; g + scalars
        NES_TRY(upload(c, bt->d_mode, mode.data(), B * sizeof(int)));
        {
            StageTimer t(c, NES_STAGE_VECTOR);
            batch_slack_kernel<<<B, 256, 0, c->stream>>>(n, np, bt->d_mode, bt->x, bt->l, bt->u, bt->c, bt->slack,
                                                        bt->theta, bt->sc, bt->d_scal);
            NES_CHECK_LAUNCH(c);
        }
        {
            StageTimer t(c, NES_STAGE_GEMV);
            batch_gemv_n_kernel<<<gm, 256, 0, c->stream>>>(bt->d_A, bt->ld, m, n, mp, np, bt->sc, bt->slack, 1.0, 0.0,
                                                          bt->t);
            NES_CHECK_LAUNCH(c);
            batch_select_rhs_kernel<<<gm, 256, 0, c->stream>>>(m, mp, bt->d_mode, bt->r, bt->t);
            NES_CHECK_LAUNCH(c);
        }
        NES_TRY(batch_factor(c, bt));
        NES_TRY(batch_solve(c, bt, bt->t));
        {
            StageTimer t(c, NES_STAGE_GEMV);
            // w = AD' t  (into theta's slot is not possible: theta is read by nobody after the factor)
            batch_gemv_t_kernel<<<gt, 256, 0, c->stream>>>(bt->d_A, bt->ld, m, n, mp, np, bt->t, bt->slack, 1.0, 0.0,
                                                          bt->theta);
            NES_CHECK_LAUNCH(c);
        }
        {
            StageTimer t(c, NES_STAGE_VECTOR);
            batch_direction_kernel<<<B, 256, 0, c->stream>>>(n, np, bt->d_mode, bt->theta, bt->sc, bt->slack, bt->x,
                                                            bt->l, bt->u, bt->c, bt->g, bt->d_scal);
            NES_CHECK_LAUNCH(c);
        }
        NES_TRY(download(c, scal.data(), bt->d_scal, scal.size() * sizeof(double)));
        NES_TRY(download(c, info.data(), bt->d_info, info.size() * sizeof(int)));
        return 0;
    };
    auto apply = [&]() -> int {
        NES_TRY(upload(c, bt->d_step, step.data(), B * sizeof(double)));
        StageTimer t(c, NES_STAGE_VECTOR);
        batch_axpy_kernel<<<gn, 256, 0, c->stream>>>(n, np, bt->d_step, bt->g, bt->x);
        NES_CHECK_LAUNCH(c);
        return 0;
    };
    int i = 0;
    for (; max_iter <= 0 || i <= max_iter; ++i) {
        // residual at the top of iteration i == residual after iteration i-1 (:284-287)
        NES_CUDA(c, cudaMemcpyAsync(bt->r, bt->b, (size_t)B * mp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        {
            StageTimer t(c, NES_STAGE_GEMV);
            batch_gemv_n_kernel<<<gm, 256, 0, c->stream>>>(bt->d_A, bt->ld, m, n, mp, np, bt->x, nullptr, -1.0, 1.0,
                                                          bt->r);
            NES_CHECK_LAUNCH(c);
        }
        {
            StageTimer t(c, NES_STAGE_VECTOR);
            batch_residual_scalars_kernel<<<B, 256, 0, c->stream>>>(m, n, mp, np, bt->r, bt->x, bt->c, bt->d_scal);
            NES_CHECK_LAUNCH(c);
        }
        NES_TRY(download(c, scal.data(), bt->d_scal, scal.size() * sizeof(double)));
        int active = 0;
        const bool centering = ((i + 1) % 16) == 0;
        for (int b = 0; b < B; ++b) {
            const double norm = std::sqrt(scal[(size_t)b * 8]);
            mode[b] = BM_SKIP;
            if (done[b]) continue;
            objv[b] = scal[(size_t)b * 8 + 1];
            resv[b] = norm;
            if (i > 0 && !(cont[b] || norm > thr)) {
                done[b] = 1;
                it_count[b] = i;
                continue;
            }
            if (max_iter > 0 && i >= max_iter) {
                done[b] = 1;
                it_count[b] = i;
                continue;
            }
            mode[b] = (norm > thr) ? BM_REPAIR : (centering ? BM_CENTER : BM_OPT);
            ++active;
        }
        if (active == 0) break;
        NES_TRY(direction_pass());
        bool redo = false;
        std::vector<int> redo_mode(B, BM_SKIP);
        auto decide = [&](int b, int md) {
            step[b] = 0.0;
            if (md == BM_SKIP) return;
            const double* s = &scal[(size_t)b * 8];
            if (info[2 * b] != 0) {  // " singular ": stop this LP
                cont[b] = 0;
                if (md == BM_REPAIR) { done[b] = 1; it_count[b] = -(i + 1); }
                return;
            }
            const double maxstep = s[2], norm_g = std::sqrt(s[3]), norm_dg = std::sqrt(s[4]), descent = s[5];
            if (md == BM_REPAIR) {
                step[b] = gamma * std::fmin(maxstep, 1.0 / gamma);
                cont[b] = 1;
                return;
            }
            const double st = gamma * maxstep;
            if (md == BM_OPT) {
                if (norm_dg < std::fmin(1e-6, 1e-8 * (double)n) || descent > 0.0) {
                    cont[b] = 0;
                    return;
                }
                if (st * norm_g < 1e-6 || descent > 0.0) {  // redo this iteration as centering (:200-204)
                    redo_mode[b] = BM_CENTER;
                    redo = true;
                    return;
                }
            }
            step[b] = st;
            cont[b] = 1;
        };
        for (int b = 0; b < B; ++b) decide(b, mode[b]);
        NES_TRY(apply());
        if (redo) {
            mode = redo_mode;
            NES_TRY(direction_pass());
            for (int b = 0; b < B; ++b) decide(b, mode[b]);
            NES_TRY(apply());
        }
    }
    for (int b = 0; b < B; ++b) {
        if (iters) iters[b] = done[b] ? it_count[b] : i;
        if (obj) obj[b] = objv[b];
        if (res) res[b] = resv[b];
    }
    return 0;
}

int nes_batch_get_x(nes_batch* bt, double* x_all, nes_ctx* c) {
    NES_ENTER(c);
    if (!bt || !x_all) return fail(c, NES_ERR_INVALID, "nes_batch_get_x: null argument");
    NES_CUDA(c, cudaMemcpy2DAsync(x_all, bt->n * sizeof(double), bt->x, bt->np * sizeof(double),
                                  bt->n * sizeof(double), bt->B, cudaMemcpyDeviceToHost, c->stream));
    NES_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
