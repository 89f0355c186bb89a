// Constraint-matrix objects: device-resident A (dense column-major or CSC) with a lazily applied
// column scale.  Replaces the cholmod_sparse*/cholmod_dense* plumbing of sparse-cholesky.lisp:
//   make-dense-from-matlisp + cholmod_dense_to_sparse   (:346-368, :411-414)
//   make-sparse-from-triplet-vector                      (:433-459)
//   cholmod_copy_sparse / cholmod_free_sparse            (:128, :75)
//   scale-sparse! = cholmod_scale(CHOLMOD_COL)           (:461-473)
// The reference copies A and rescales the copy's values every iteration; here the copy shares
// the immutable values and only stores s, which the formation and matvec kernels apply on the fly.
#include <algorithm>
#include <numeric>

#include "dmma_nt.cuh"
#include "nes_internal.h"

using namespace nes;

namespace nes {

// counter-based generator shared with lpgen.dense_entry (host, NumPy): splitmix64 finaliser
__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__global__ void generate_dense_kernel(double* A, size_t m, size_t n, size_t ld,
                                      unsigned long long seed) {
    const size_t total = ld * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx / ld, i = idx - j * ld;
        double v = 0.0;
        if (i < m) {
            unsigned long long h = mix64(mix64(seed) ^ (i * 0x100000001B3ULL + j * 0xC2B2AE3D27D4EB4FULL));
            v = (double)(h >> 11) * (1.0 / 9007199254740992.0);
            if (i == j) v += 1.0;
        }
        A[idx] = v;
    }
}

__global__ void square_pad_kernel(const double* s, double* theta, size_t n, size_t npad) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < npad) theta[i] = (i < n) ? s[i] * s[i] : 0.0;
}

// one warp per row block: max |a_ij| per row of a dense column-major matrix, then scale rows
__global__ void row_maxabs_dense_kernel(const double* A, size_t m, size_t n, size_t ld,
                                        double* rowmax) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    double mx = 0.0;
    for (size_t j = 0; j < n; ++j) mx = fmax(mx, fabs(A[i + j * ld]));
    rowmax[i] = (mx < 1e-6) ? 1.0 : 1.0 / mx;
}

__global__ void scale_rows_dense_kernel(double* A, size_t m, size_t n, size_t ld,
                                        const double* rowscale) {
    const size_t total = ld * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx / ld, i = idx - j * ld;
        if (i < m) A[idx] *= rowscale[i];
    }
}

__global__ void row_maxabs_csc_kernel(const int* rowidx, const double* val, size_t nnz,
                                      unsigned long long* rowmax_bits) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    // non-negative doubles order like their bit patterns
    atomicMax(&rowmax_bits[rowidx[k]], (unsigned long long)__double_as_longlong(fabs(val[k])));
}

__global__ void rowmax_to_scale_kernel(const unsigned long long* bits, double* scale, size_t m) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    double mx = __longlong_as_double((long long)bits[i]);
    // rows without entries are left alone (the Lisp only visits rows that appear in a triplet)
    scale[i] = (mx < 1e-6) ? 1.0 : 1.0 / mx;
}

__global__ void scale_rows_csc_kernel(const int* rowidx, double* val, size_t nnz,
                                      const double* rowscale) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < nnz) val[k] *= rowscale[rowidx[k]];
}

__global__ void gather_kernel(const double* __restrict__ src, const int* __restrict__ idx,
                              double* __restrict__ dst, size_t n) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < n) dst[k] = src[idx[k]];
}

static void free_base(nes_ctx* c, MatrixBase* b) {
    if (!b) return;
    if (--b->refs > 0) return;
    dev_free(c, b->d_val);
    dev_free(c, b->d_colptr);
    dev_free(c, b->d_rowidx);
    dev_free(c, b->d_values);
    dev_free(c, b->d_rowptr);
    dev_free(c, b->d_colidx);
    dev_free(c, b->d_csr_src);
    dev_free(c, b->d_csr_val);
    delete b;
}

static int refresh_csr_values(nes_ctx* c, MatrixBase* b) {
    if (b->nnz == 0) return 0;
    gather_kernel<<<(unsigned)((b->nnz + 255) / 256), 256, 0, c->stream>>>(b->d_values, b->d_csr_src,
                                                                          b->d_csr_val, b->nnz);
    NES_CHECK_LAUNCH(c);
    return 0;
}

static nes_matrix* new_dense(nes_ctx* c, size_t m, size_t n) {
    MatrixBase* b = new MatrixBase();
    b->dense = true;
    b->m = m;
    b->n = n;
    b->ld = (m + 15) / 16 * 16;
    if (b->ld == 0) b->ld = 16;
    b->d_val = static_cast<double*>(dev_alloc(c, b->ld * (n ? n : 1) * sizeof(double)));
    if (!b->d_val) {
        delete b;
        return nullptr;
    }
    if (make_operand_map(&b->map, b->d_val, (long long)m, (long long)n, (long long)b->ld) != 0) {
        fail(c, NES_ERR_CUDA, "cuTensorMapEncodeTiled failed for A (%zu x %zu)", m, n);
        dev_free(c, b->d_val);
        delete b;
        return nullptr;
    }
    nes_matrix* A = new nes_matrix();
    A->set_base(b);
    return A;
}

}  // namespace nes

extern "C" {

nes_matrix* nes_dense_to_matrix(const double* Ah, size_t nrow, size_t ncol, size_t ld, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (!Ah || ld < nrow || nrow == 0 || ncol == 0) {
        fail(c, NES_ERR_INVALID, "nes_dense_to_matrix: bad arguments");
        return nullptr;
    }
    nes_matrix* A = new_dense(c, nrow, ncol);
    if (!A) return nullptr;
    MatrixBase* b = A->base;
    cudaError_t e = cudaMemsetAsync(b->d_val, 0, b->ld * ncol * sizeof(double), c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(b->d_val, b->ld * sizeof(double), Ah, ld * sizeof(double),
                              nrow * sizeof(double), ncol, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        fail(c, NES_ERR_CUDA, "upload of A failed: %s", cudaGetErrorString(e));
        nes_free_matrix(&A, c);
        return nullptr;
    }
    return A;
}

nes_matrix* nes_generate_dense(size_t nrow, size_t ncol, unsigned long long seed, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (nrow == 0 || ncol == 0) {
        fail(c, NES_ERR_INVALID, "nes_generate_dense: empty matrix");
        return nullptr;
    }
    nes_matrix* A = new_dense(c, nrow, ncol);
    if (!A) return nullptr;
    MatrixBase* b = A->base;
    generate_dense_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(b->d_val, nrow, ncol, b->ld, seed);
    ++c->launches;
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
        fail(c, NES_ERR_CUDA, "generate_dense_kernel failed");
        nes_free_matrix(&A, c);
        return nullptr;
    }
    return A;
}

nes_matrix* nes_csc_to_matrix(const int* colptr, const int* rowidx, const double* val, size_t nrow,
                              size_t ncol, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (!colptr || (colptr[ncol] > 0 && (!rowidx || !val))) {
        fail(c, NES_ERR_INVALID, "nes_csc_to_matrix: null input");
        return nullptr;
    }
    const size_t nnz_in = (size_t)colptr[ncol];
    // sort rows inside each column and sum duplicates (cholmod_triplet_to_sparse + cholmod_sort)
    std::vector<int> cp(ncol + 1, 0), ri;
    std::vector<double> vx;
    ri.reserve(nnz_in);
    vx.reserve(nnz_in);
    std::vector<std::pair<int, double>> colbuf;
    for (size_t j = 0; j < ncol; ++j) {
        colbuf.clear();
        for (int k = colptr[j]; k < colptr[j + 1]; ++k) {
            if (rowidx[k] < 0 || (size_t)rowidx[k] >= nrow) {
                fail(c, NES_ERR_INVALID, "row index %d out of range in column %zu", rowidx[k], j);
                return nullptr;
            }
            colbuf.emplace_back(rowidx[k], val[k]);
        }
        std::stable_sort(colbuf.begin(), colbuf.end(),
                         [](const std::pair<int, double>& a, const std::pair<int, double>& b) {
                             return a.first < b.first;
                         });
        for (size_t k = 0; k < colbuf.size(); ++k) {
            if (!ri.empty() && (int)ri.size() > cp[j] && ri.back() == colbuf[k].first)
                vx.back() += colbuf[k].second;
            else {
                ri.push_back(colbuf[k].first);
                vx.push_back(colbuf[k].second);
            }
        }
        cp[j + 1] = (int)ri.size();
    }
    MatrixBase* b = new MatrixBase();
    b->dense = false;
    b->m = nrow;
    b->n = ncol;
    b->nnz = ri.size();
    b->d_colptr = static_cast<int*>(dev_alloc(c, (ncol + 1) * sizeof(int)));
    b->d_rowidx = static_cast<int*>(dev_alloc(c, b->nnz * sizeof(int)));
    b->d_values = static_cast<double*>(dev_alloc(c, b->nnz * sizeof(double)));
    if (!b->d_colptr || !b->d_rowidx || !b->d_values ||
        upload(c, b->d_colptr, cp.data(), (ncol + 1) * sizeof(int)) != 0 ||
        upload(c, b->d_rowidx, ri.data(), b->nnz * sizeof(int)) != 0 ||
        upload(c, b->d_values, vx.data(), b->nnz * sizeof(double)) != 0) {
        free_base(c, b);
        return nullptr;
    }
    // CSR mirror for the atomic-free row-gather product
    {
        std::vector<int> rp(nrow + 1, 0), cj(b->nnz), src(b->nnz);
        for (size_t k = 0; k < b->nnz; ++k) rp[ri[k] + 1]++;
        for (size_t i = 0; i < nrow; ++i) rp[i + 1] += rp[i];
        std::vector<int> next(rp.begin(), rp.end() - 1);
        for (size_t j = 0; j < ncol; ++j)
            for (int k = cp[j]; k < cp[j + 1]; ++k) {
                const int q = next[ri[k]]++;
                cj[q] = (int)j;
                src[q] = k;
            }
        b->d_rowptr = static_cast<int*>(dev_alloc(c, (nrow + 1) * sizeof(int)));
        b->d_colidx = static_cast<int*>(dev_alloc(c, b->nnz * sizeof(int)));
        b->d_csr_src = static_cast<int*>(dev_alloc(c, b->nnz * sizeof(int)));
        b->d_csr_val = static_cast<double*>(dev_alloc(c, b->nnz * sizeof(double)));
        if (!b->d_rowptr || !b->d_colidx || !b->d_csr_src || !b->d_csr_val ||
            upload(c, b->d_rowptr, rp.data(), (nrow + 1) * sizeof(int)) != 0 ||
            upload(c, b->d_colidx, cj.data(), b->nnz * sizeof(int)) != 0 ||
            upload(c, b->d_csr_src, src.data(), b->nnz * sizeof(int)) != 0 ||
            refresh_csr_values(c, b) != 0) {
            free_base(c, b);
            return nullptr;
        }
    }
    b->h_colptr = std::move(cp);
    b->h_rowidx = std::move(ri);
    nes_matrix* A = new nes_matrix();
    A->set_base(b);
    return A;
}

nes_matrix* nes_triplet_to_sparse(const int* row, const int* col, const double* val, size_t nnz,
                                  size_t nrow, size_t ncol, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (nnz > 0 && (!row || !col || !val)) {
        fail(c, NES_ERR_INVALID, "nes_triplet_to_sparse: null input");
        return nullptr;
    }
    std::vector<int> cp(ncol + 1, 0), ri(nnz);
    std::vector<double> vx(nnz);
    for (size_t k = 0; k < nnz; ++k) {
        // asserts of make-sparse-from-triplet-vector (sparse-cholesky.lisp:450-451)
        if (row[k] < 0 || (size_t)row[k] >= nrow || col[k] < 0 || (size_t)col[k] >= ncol) {
            fail(c, NES_ERR_INVALID, "triplet %zu (%d,%d) out of range", k, row[k], col[k]);
            return nullptr;
        }
        cp[col[k] + 1]++;
    }
    for (size_t j = 0; j < ncol; ++j) cp[j + 1] += cp[j];
    std::vector<int> next(cp.begin(), cp.end() - 1);
    for (size_t k = 0; k < nnz; ++k) {
        int p = next[col[k]]++;
        ri[p] = row[k];
        vx[p] = val[k];
    }
    return nes_csc_to_matrix(cp.data(), ri.data(), vx.data(), nrow, ncol, c);
}

nes_matrix* nes_copy_matrix(nes_matrix* A, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (!A) {
        fail(c, NES_ERR_INVALID, "nes_copy_matrix: null");
        return nullptr;
    }
    nes_matrix* B = new nes_matrix();
    B->set_base(A->base);
    B->base->refs++;
    if (A->d_scale) {
        const size_t n = A->base->n, npad = (n + 15) / 16 * 16;
        B->d_scale = static_cast<double*>(dev_alloc(c, n * sizeof(double)));
        B->d_theta = static_cast<double*>(dev_alloc(c, npad * sizeof(double)));
        if (!B->d_scale || !B->d_theta) {
            nes_free_matrix(&B, c);
            return nullptr;
        }
        cudaMemcpyAsync(B->d_scale, A->d_scale, n * sizeof(double), cudaMemcpyDeviceToDevice,
                        c->stream);
        cudaMemcpyAsync(B->d_theta, A->d_theta, npad * sizeof(double), cudaMemcpyDeviceToDevice,
                        c->stream);
    }
    return B;
}

int nes_free_matrix(nes_matrix** A, nes_ctx* c) {
    if (!c) return 0;
    if (!A || !*A) return 1;  // idempotent on NULL
    if (c->started) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
    }
    dev_free(c, (*A)->d_scale);
    dev_free(c, (*A)->d_theta);
    free_base(c, (*A)->base);
    delete *A;
    *A = nullptr;
    return 1;
}

}  // extern "C"

namespace nes {
// device-side entry used by the IPM drivers: s already lives on the GPU
int set_scale_dev(nes_ctx* c, nes_matrix* A, const double* d_s) {
    const size_t n = A->base->n, npad = (n + 15) / 16 * 16;
    if (!A->d_scale) {
        A->d_scale = static_cast<double*>(dev_alloc(c, n * sizeof(double)));
        A->d_theta = static_cast<double*>(dev_alloc(c, npad * sizeof(double)));
        if (!A->d_scale || !A->d_theta) return c->status;
    }
    if (d_s != A->d_scale)
        NES_CUDA(c, cudaMemcpyAsync(A->d_scale, d_s, n * sizeof(double), cudaMemcpyDeviceToDevice,
                                    c->stream));
    square_pad_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(A->d_scale, A->d_theta,
                                                                             n, npad);
    NES_CHECK_LAUNCH(c);
    return 0;
}
}  // namespace nes

extern "C" {

int nes_scale(const double* s, int scale, nes_matrix* A, nes_ctx* c) {
    NES_ENTER(c);
    if (!A || !s) return fail(c, NES_ERR_INVALID, "nes_scale: null argument"), 0;
    if (scale != 2) return fail(c, NES_ERR_INVALID, "nes_scale: only CHOLMOD_COL (2) is supported"), 0;
    const size_t n = A->base->n;
    const bool rescale = (A->d_scale != nullptr);
    double* d_tmp = ensure_ws(c, WS_API, n * sizeof(double));
    if (!d_tmp) return 0;
    if (upload(c, d_tmp, s, n * sizeof(double)) != 0) return 0;
    if (rescale) {
        // cholmod_scale on an already scaled matrix compounds the factors
        // (only reached if a caller scales twice without nes_unscale).
        std::vector<double> old(n), neu(n);
        if (download(c, old.data(), A->d_scale, n * sizeof(double)) != 0) return 0;
        for (size_t i = 0; i < n; ++i) neu[i] = old[i] * s[i];
        if (upload(c, d_tmp, neu.data(), n * sizeof(double)) != 0) return 0;
    }
    if (set_scale_dev(c, A, d_tmp) != 0) return 0;
    return 1;  // CHOLMOD: TRUE on success (asserted non-zero at sparse-cholesky.lisp:468-471)
}

int nes_unscale(nes_matrix* A, nes_ctx* c) {
    NES_ENTER(c);
    if (!A) return 0;
    cudaStreamSynchronize(c->stream);
    dev_free(c, A->d_scale);
    dev_free(c, A->d_theta);
    A->d_scale = nullptr;
    A->d_theta = nullptr;
    return 1;
}

size_t nes_matrix_nrow(const nes_matrix* A) { return A ? A->base->m : 0; }
size_t nes_matrix_ncol(const nes_matrix* A) { return A ? A->base->n : 0; }
size_t nes_matrix_nnz(const nes_matrix* A) {
    if (!A) return 0;
    return A->base->dense ? A->base->m * A->base->n : A->base->nnz;
}
int nes_matrix_is_dense(const nes_matrix* A) { return A && A->base->dense ? 1 : 0; }

int nes_scale_rows_maxabs(nes_matrix* A, double* rowscale_out, nes_ctx* c) {
    NES_ENTER(c);
    if (!A) return fail(c, NES_ERR_INVALID, "nes_scale_rows_maxabs: null");
    MatrixBase* b = A->base;
    const size_t m = b->m;
    double* d_rs = ensure_ws(c, WS_API, 2 * m * sizeof(double));
    if (!d_rs) return c->status;
    if (b->dense) {
        row_maxabs_dense_kernel<<<(unsigned)((m + 127) / 128), 128, 0, c->stream>>>(b->d_val, m, b->n,
                                                                                   b->ld, d_rs);
        NES_CHECK_LAUNCH(c);
        scale_rows_dense_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(b->d_val, m, b->n, b->ld, d_rs);
        NES_CHECK_LAUNCH(c);
    } else {
        unsigned long long* bits = reinterpret_cast<unsigned long long*>(d_rs + m);
        NES_CUDA(c, cudaMemsetAsync(bits, 0, m * sizeof(unsigned long long), c->stream));
        if (b->nnz) {
            row_maxabs_csc_kernel<<<(unsigned)((b->nnz + 255) / 256), 256, 0, c->stream>>>(
                b->d_rowidx, b->d_values, b->nnz, bits);
            NES_CHECK_LAUNCH(c);
        }
        rowmax_to_scale_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(bits, d_rs, m);
        NES_CHECK_LAUNCH(c);
        if (b->nnz) {
            scale_rows_csc_kernel<<<(unsigned)((b->nnz + 255) / 256), 256, 0, c->stream>>>(
                b->d_rowidx, b->d_values, b->nnz, d_rs);
            NES_CHECK_LAUNCH(c);
            NES_TRY(refresh_csr_values(c, b));
        }
    }
    if (rowscale_out) NES_TRY(download(c, rowscale_out, d_rs, m * sizeof(double)));
    return 0;
}

}  // extern "C"
