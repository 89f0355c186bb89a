// K6: y <- alpha * op(A diag(s)) x + beta * y on device vectors.
// Replaces cholmod_sdmult as driven by sparse-m* (sparse-cholesky.lisp:567-614) and the dense
// gemm!/m* matrix-vector products of newton-solve.lisp:62, 98, 129.  All variants are HBM-bound
// (dense: 8mn bytes per product) and atomic-free, so results are bitwise reproducible.
//   dense  N : row-pair per thread (16B loads), columns split across CTAs, partials + fixed-order sum
//   dense  T : one warp per column pair, 16B loads along the column, shuffle reduction
//   CSC    N : row gather over the CSR mirror (1-8 lanes per row, fixed shuffle tree)
//   CSC    T : column gather (1-8 lanes per column)
// The column scale s of nes_scale is folded in: N uses x_k*s_k, T multiplies the result by s_j.
#include <cstdlib>

#include "nes_internal.h"

namespace nes {

constexpr int GN_THREADS = 128;   // rows per CTA = 256
constexpr int GN_CHUNK = 512;     // columns staged per smem refill

__global__ void __launch_bounds__(GN_THREADS)
gemv_n_partial_kernel(const double* __restrict__ A, size_t ld, int m, int n, int cols_per_cta,
                      const double* __restrict__ x, const double* __restrict__ s,
                      double* __restrict__ partial, size_t ldp) {
    __shared__ double xs[GN_CHUNK];
    const int tid = threadIdx.x;
    const int r = (blockIdx.x * GN_THREADS + tid) * 2;
    const int k_begin = blockIdx.y * cols_per_cta;
    const int k_end = min(n, k_begin + cols_per_cta);
    double a0 = 0.0, a1 = 0.0;
    for (int kc = k_begin; kc < k_end; kc += GN_CHUNK) {
        const int kn = min(GN_CHUNK, k_end - kc);
        __syncthreads();
        for (int i = tid; i < kn; i += GN_THREADS) xs[i] = s ? x[kc + i] * s[kc + i] : x[kc + i];
        __syncthreads();
        if (r < m) {
            const double* Ap = A + r + (size_t)kc * ld;
#pragma unroll 8
            for (int k = 0; k < kn; ++k) {
                const double2 v = *reinterpret_cast<const double2*>(Ap + (size_t)k * ld);
                a0 = fma(v.x, xs[k], a0);
                a1 = fma(v.y, xs[k], a1);
            }
        }
    }
    if (r < m) {
        partial[blockIdx.y * ldp + r] = a0;
        if (r + 1 < m) partial[blockIdx.y * ldp + r + 1] = a1;
    }
}

__global__ void gemv_n_finish_kernel(const double* __restrict__ partial, size_t ldp, int nparts, int m,
                                     double alpha, double beta, double* __restrict__ y) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += partial[p * ldp + r];
    y[r] = (beta == 0.0) ? alpha * acc : fma(alpha, acc, beta * y[r]);
}

// warp handles columns j, j+1
__global__ void __launch_bounds__(256)
gemv_t_kernel(const double* __restrict__ A, size_t ld, int m, int n, const double* __restrict__ x,
              const double* __restrict__ s, double alpha, double beta, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 2;
    if (j >= n) return;
    const bool two = (j + 1 < n);
    const double* A0 = A + (size_t)j * ld;
    const double* A1 = A0 + (two ? ld : 0);
    double acc0 = 0.0, acc1 = 0.0;
    const int m2 = m & ~1;
#pragma unroll 4
    for (int i = lane * 2; i < m2; i += 64) {
        const double2 xv = *reinterpret_cast<const double2*>(x + i);
        const double2 v0 = *reinterpret_cast<const double2*>(A0 + i);
        const double2 v1 = *reinterpret_cast<const double2*>(A1 + i);
        acc0 = fma(v0.x, xv.x, acc0);
        acc0 = fma(v0.y, xv.y, acc0);
        acc1 = fma(v1.x, xv.x, acc1);
        acc1 = fma(v1.y, xv.y, acc1);
    }
    if (lane == 0 && m2 < m) {
        acc0 = fma(A0[m2], x[m2], acc0);
        acc1 = fma(A1[m2], x[m2], acc1);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
    }
    if (lane == 0) {
        if (s) acc0 *= s[j];
        y[j] = (beta == 0.0) ? alpha * acc0 : fma(alpha, acc0, beta * y[j]);
        if (two) {
            if (s) acc1 *= s[j + 1];
            y[j + 1] = (beta == 0.0) ? alpha * acc1 : fma(alpha, acc1, beta * y[j + 1]);
        }
    }
}

__global__ void axpby_kernel(int n, double alpha, const double* __restrict__ t, double beta,
                             double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (beta == 0.0) ? alpha * t[i] : fma(alpha, t[i], beta * y[i]);
}

int dist_allreduce_sum(nes_ctx* c, double* d_buf, size_t count);  // nes_dist.cu

// Multi-GPU dense products: A is replicated, so every rank multiplies only ITS block
// of columns -- 1/Q of the HBM traffic -- and one all-reduce of the m-vector (forward) or of the
// zero-padded n-vector (transposed) completes the product; NCCL's all-reduce returns bit-identical
// results on every rank, so the replicated control flow of the IPM stays in lockstep.
// Sub-warp CSR / CSC gathers: LANES threads share one row (column), reading consecutive entries (coalesced
// index + value segments), and combine with a fixed shuffle tree -- deterministic, and 2-3x the bandwidth of
// one thread per row once rows hold more than a few entries.
template <int LANES>
__global__ void __launch_bounds__(256)
spmv_gather_kernel(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                   int nrows, const double* __restrict__ x, const double* __restrict__ s_in,
                   const double* __restrict__ s_out, double alpha, double beta, double* __restrict__ y) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = t / LANES, lane = t % LANES;
    double acc = 0.0;
    if (r < nrows) {
        const int e = ptr[r + 1];
        for (int k = ptr[r] + lane; k < e; k += LANES) {
            const int j = idx[k];
            acc = fma(val[k], s_in ? x[j] * s_in[j] : x[j], acc);
        }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
    if (r < nrows && lane == 0) {
        if (s_out) acc *= s_out[r];
        y[r] = (beta == 0.0) ? alpha * acc : fma(alpha, acc, beta * y[r]);
    }
}

template <int LANES>
static void spmv_gather_launch(cudaStream_t st, const int* ptr, const int* idx, const double* val, int nrows,
                               const double* x, const double* s_in, const double* s_out, double alpha, double beta,
                               double* y) {
    const long long threads = (long long)nrows * LANES;
    spmv_gather_kernel<LANES><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(ptr, idx, val, nrows, x, s_in, s_out,
                                                                               alpha, beta, y);
}

static void spmv_gather(cudaStream_t st, const int* ptr, const int* idx, const double* val, int nrows, size_t nnz,
                        const double* x, const double* s_in, const double* s_out, double alpha, double beta,
                        double* y) {
    double avg = nrows > 0 ? (double)nnz / nrows : 0.0;
    static int forced = -1;
    if (forced < 0) forced = getenv("NES_SPMV_LANES") ? atoi(getenv("NES_SPMV_LANES")) : 0;
    if (forced == 32) avg = 1000.0;
    else if (forced == 16) avg = 300.0;
    else if (forced == 8) avg = 100.0;
    else if (forced == 4) avg = 6.0;
    else if (forced == 2) avg = 3.0;
    else if (forced == 1) avg = 0.0;
    if (avg >= 500.0) spmv_gather_launch<32>(st, ptr, idx, val, nrows, x, s_in, s_out, alpha, beta, y);
    else if (avg >= 200.0) spmv_gather_launch<16>(st, ptr, idx, val, nrows, x, s_in, s_out, alpha, beta, y);
    else if (avg >= 12.0) spmv_gather_launch<8>(st, ptr, idx, val, nrows, x, s_in, s_out, alpha, beta, y);
    else if (avg >= 5.0) spmv_gather_launch<4>(st, ptr, idx, val, nrows, x, s_in, s_out, alpha, beta, y);
    else if (avg >= 2.5) spmv_gather_launch<2>(st, ptr, idx, val, nrows, x, s_in, s_out, alpha, beta, y);
    else spmv_gather_launch<1>(st, ptr, idx, val, nrows, x, s_in, s_out, alpha, beta, y);
}

static int matvec_dense_split(nes_ctx* c, const MatrixBase* b, const double* d_s, int transpose, double alpha,
                              const double* d_x, double beta, double* d_y) {
    const int m = (int)b->m, n = (int)b->n, Q = c->nranks, r = c->rank;
    const int k0 = (int)((long long)n * r / Q) & ~15, k1 = (r + 1 == Q) ? n : ((int)((long long)n * (r + 1) / Q) & ~15);
    const int nsub = k1 - k0;
    const double* Asub = b->d_val + (size_t)k0 * b->ld;
    const double* ssub = d_s ? d_s + k0 : nullptr;
    if (!transpose) {
        const int rowblocks = (m + 2 * GN_THREADS - 1) / (2 * GN_THREADS);
        int want = (c->num_sms * 4 + rowblocks - 1) / rowblocks;
        int cols = (nsub + want - 1) / want;
        cols = (cols + 63) / 64 * 64;
        if (cols < 64) cols = 64;
        const int nparts = nsub > 0 ? (nsub + cols - 1) / cols : 0;
        const size_t ldp = (size_t)(m + 1) / 2 * 2;
        double* ws = ensure_ws(c, WS_MATVEC, ((size_t)(nparts + 1) * ldp + 16) * sizeof(double));
        if (!ws) return c->status;
        double* sum = ws;
        double* part = ws + ldp;
        if (nparts > 0) {
            gemv_n_partial_kernel<<<dim3(rowblocks, nparts), GN_THREADS, 0, c->stream>>>(
                Asub, b->ld, m, nsub, cols, d_x + k0, ssub, part, ldp);
            NES_CHECK_LAUNCH(c);
        }
        gemv_n_finish_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(part, ldp, nparts, m, 1.0, 0.0, sum);
        NES_CHECK_LAUNCH(c);
        NES_TRY(dist_allreduce_sum(c, sum, (size_t)m));
        axpby_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(m, alpha, sum, beta, d_y);
        NES_CHECK_LAUNCH(c);
    } else {
        double* tmp = ensure_ws(c, WS_MATVEC, ((size_t)n + 16) * sizeof(double));
        if (!tmp) return c->status;
        NES_CUDA(c, cudaMemsetAsync(tmp, 0, (size_t)n * sizeof(double), c->stream));
        if (nsub > 0) {
            gemv_t_kernel<<<(nsub + 15) / 16, 256, 0, c->stream>>>(Asub, b->ld, m, nsub, d_x, ssub, 1.0, 0.0,
                                                                   tmp + k0);
            NES_CHECK_LAUNCH(c);
        }
        NES_TRY(dist_allreduce_sum(c, tmp, (size_t)n));
        axpby_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(n, alpha, tmp, beta, d_y);
        NES_CHECK_LAUNCH(c);
    }
    return 0;
}

static int matvec_impl(nes_ctx* c, const MatrixBase* b, const double* d_s, int transpose, double alpha,
                       const double* d_x, double beta, double* d_y) {
    StageTimer timer(c, NES_STAGE_GEMV);
    const int m = (int)b->m, n = (int)b->n;
    if (b->dense && c->nranks > 1 && c->nccl_comm && (long long)m * n >= (1 << 18) && n >= 64 * c->nranks)
        return matvec_dense_split(c, b, d_s, transpose, alpha, d_x, beta, d_y);
    if (b->dense) {
        if (!transpose) {
            const int rowblocks = (m + 2 * GN_THREADS - 1) / (2 * GN_THREADS);
            int want = (c->num_sms * 4 + rowblocks - 1) / rowblocks;
            int cols = (n + want - 1) / want;
            cols = (cols + 63) / 64 * 64;
            const int nparts = (n + cols - 1) / cols;
            const size_t ldp = (size_t)(m + 1) / 2 * 2;
            double* part = ensure_ws(c, WS_MATVEC, (size_t)nparts * ldp * sizeof(double));
            if (!part) return c->status;
            gemv_n_partial_kernel<<<dim3(rowblocks, nparts), GN_THREADS, 0, c->stream>>>(
                b->d_val, b->ld, m, n, cols, d_x, d_s, part, ldp);
            NES_CHECK_LAUNCH(c);
            gemv_n_finish_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(part, ldp, nparts, m, alpha,
                                                                        beta, d_y);
            NES_CHECK_LAUNCH(c);
        } else {
            gemv_t_kernel<<<(n + 15) / 16, 256, 0, c->stream>>>(b->d_val, b->ld, m, n, d_x, d_s, alpha,
                                                                beta, d_y);
            NES_CHECK_LAUNCH(c);
        }
    } else {
        if (!transpose) {  // rows of the CSR mirror; the column scale multiplies x
            spmv_gather(c->stream, b->d_rowptr, b->d_colidx, b->d_csr_val, m, b->nnz, d_x, d_s, nullptr, alpha,
                        beta, d_y);
            NES_CHECK_LAUNCH(c);
        } else {           // columns of the CSC arrays; the column scale multiplies the result
            spmv_gather(c->stream, b->d_colptr, b->d_rowidx, b->d_values, n, b->nnz, d_x, nullptr, d_s, alpha,
                        beta, d_y);
            NES_CHECK_LAUNCH(c);
        }
    }
    return 0;
}

int matvec(nes_ctx* c, const nes_matrix* A, int transpose, double alpha, const double* d_x,
           double beta, double* d_y) {
    return matvec_impl(c, A->base, A->d_scale, transpose, alpha, d_x, beta, d_y);
}

int matvec_unscaled(nes_ctx* c, const MatrixBase* A, int transpose, double alpha, const double* d_x,
                    double beta, double* d_y) {
    return matvec_impl(c, A, nullptr, transpose, alpha, d_x, beta, d_y);
}

}  // namespace nes
