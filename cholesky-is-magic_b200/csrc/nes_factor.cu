// analyze / factorize / solve / sdmult entry points: the CHOLMOD-shaped part of the boundary
// (sparse-cholesky.lisp:261-288, 335-342 declarations; call sites :409-431, :506-560, :567-614).
#include <cmath>
#include <vector>

#include "dmma_nt.cuh"
#include "nes_internal.h"

using namespace nes;

namespace nes {
int sparse_analyze(nes_ctx* c, nes_matrix* A, nes_factor* L);       // sparse_chol.cu
int sparse_factorize(nes_ctx* c, nes_matrix* A, nes_factor* L);
int sparse_solve_inplace(nes_ctx* c, nes_factor* L, double* d_x);
void sparse_free(nes_ctx* c, nes_factor* L);
int sparse_factor_to_dense(nes_ctx* c, nes_factor* L, double* Lout, size_t ld, int* perm_out);

int factorize_dev(nes_ctx* c, nes_matrix* A, nes_factor* L) {
    c->status = 0;  // the Lisp does cholmod_set_status 0 before every factorize (:418, :511, :541)
    if (!L->dense) return sparse_factorize(c, A, L);
    L->factorized = 0;
    NES_TRY(dense_form_normal(c, A, L, true));
    return dense_cholesky(c, L, A);
}

// ---- device-side check of the north-star gate ||L L' - M||_F / ||M||_F on the WHOLE matrix -------------
// Lc = tril(L) with explicit zeros above the diagonal (the factor's buffer keeps stale values there)
__global__ void tril_copy_kernel(const double* __restrict__ L, double* __restrict__ out, long long ld, int m) {
    const long long total = ld * (long long)m;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx / ld), i = (int)(idx - (long long)j * ld);
        out[idx] = (i >= j && i < m) ? L[idx] : 0.0;
    }
}

// sum of squares over the lower triangle, off-diagonal entries counted twice (= the full symmetric
// matrix's Frobenius norm squared); per-block partials, combined in fixed order by the caller's finish
__global__ void frob_lower_kernel(const double* __restrict__ T, long long ld, int m, double* __restrict__ partial) {
    __shared__ double buf[32];
    double acc = 0.0;
    const long long total = ld * (long long)m;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx / ld), i = (int)(idx - (long long)j * ld);
        if (i >= j && i < m) {
            const double v = T[idx];
            acc = fma(i == j ? 1.0 : 2.0, v * v, acc);
        }
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += buf[w];
        partial[blockIdx.x] = v;
    }
}

int solve_dev(nes_ctx* c, nes_factor* L, double* d_x) {
    if (!L->factorized) return fail(c, NES_ERR_INVALID, "solve on a factor that is not factorized");
    if (!L->dense) return sparse_solve_inplace(c, L, d_x);
    return dense_solve_inplace(c, L, d_x);
}
}  // namespace nes

extern "C" {

nes_factor* nes_analyze(nes_matrix* A, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (!A) {
        fail(c, NES_ERR_INVALID, "nes_analyze: null matrix");
        return nullptr;
    }
    nes_factor* L = new nes_factor();
    L->analyzed_for = A;
    L->m = A->base->m;
    if (!A->base->dense) {
        L->dense = false;
        if (sparse_analyze(c, A, L) != 0) {
            nes_free_factor(&L, c);
            return nullptr;
        }
        return L;
    }
    const size_t m = L->m, n = A->base->n;
    L->dense = true;
    L->ld = (m + 15) / 16 * 16;
    L->d_M = static_cast<double*>(dev_alloc(c, L->ld * m * sizeof(double)));
    L->d_dinv = static_cast<double*>(dev_alloc(c, (m + 16) * sizeof(double)));
    L->d_rhs = static_cast<double*>(dev_alloc(c, (m + 16) * sizeof(double)));
    L->d_info = static_cast<int*>(dev_alloc(c, 4 * sizeof(int)));
    L->d_flags = static_cast<int*>(dev_alloc(c, ((m + 127) / 128 + 1) * sizeof(int)));
    if (L->d_flags) cudaMemsetAsync(L->d_flags, 0, ((m + 127) / 128 + 1) * sizeof(int), c->stream);
    L->d_Winv = static_cast<double*>(dev_alloc(c, ((m + 127) / 128) * 128 * 128 * sizeof(double)));
    if (!L->d_M || !L->d_dinv || !L->d_rhs || !L->d_info || !L->d_flags || !L->d_Winv) {
        nes_free_factor(&L, c);
        return nullptr;
    }
    // pad rows m..ld of M are never written by the kernels; keep them finite for the TMA boxes
    cudaMemsetAsync(L->d_M, 0, L->ld * m * sizeof(double), c->stream);
    if (make_operand_map(&L->mapM, L->d_M, (long long)m, (long long)m, (long long)L->ld) != 0 ||
        make_operand_map(&L->mapM68, L->d_M, (long long)m, (long long)m, (long long)L->ld, 68, NT_BK) != 0 ||
        make_operand_map(&L->mapBlk, L->d_M, (long long)m, (long long)m, (long long)L->ld, 128, 128) != 0 ||
        make_operand_map(&L->mapSlab, L->d_M, (long long)m, (long long)m, (long long)L->ld, 64, 128) != 0) {
        fail(c, NES_ERR_CUDA, "cuTensorMapEncodeTiled failed for M (%zu x %zu)", m, m);
        nes_free_factor(&L, c);
        return nullptr;
    }
    // NES_FORCE_DIST: run the distributed schedule on ONE GPU (every message is local, no NCCL call) so that
    // it can be debugged and profiled without a multi-GPU box
    if (c->nranks > 1 || getenv("NES_FORCE_DIST")) {
        // distributed factorization: message schedule + owned-tile list (nes_dist.cu); events and the staging
        // ring are created on the first factorization
        L->nbo = dense_outer_block((int)m, c->nranks);
        L->dist = new DistPlan();
        L->ntiles_owned = dist_make_plan(*L->dist, (int)m, L->nbo, c->grid_p, c->grid_q, c->rank,
                                         dist_chunk_rows((int)m, L->nbo, c->grid_p));
        L->d_tile_list = static_cast<int2*>(dev_alloc(c, (L->dist->tiles.size() + 1) * sizeof(int2)));
        if (!L->d_tile_list ||
            upload(c, L->d_tile_list, L->dist->tiles.data(), L->dist->tiles.size() * sizeof(int2)) != 0) {
            nes_free_factor(&L, c);
            return nullptr;
        }
    }
    // the analytic counters the reference prints after cholmod_analyze (affine-scaling.lisp:273-279)
    const double dm = (double)m, dn = (double)n;
    c->anz = dm * (dm + 1) / 2;
    c->aatfl = dm * (dm + 1) * dn;
    c->lnz = dm * (dm + 1) / 2;
    c->fl = dm * (dm + 1) * (2 * dm + 1) / 6;
    c->status = 0;
    return L;
}

int nes_factorize(nes_matrix* A, nes_factor* L, nes_ctx* c) {
    if (!c || !c->started) return 0;
    cudaSetDevice(c->device);
    if (!A || !L || A->base->m != L->m || A->base->dense != L->dense) {
        fail(c, NES_ERR_INVALID, "nes_factorize: factor was not analyzed for this matrix");
        return 0;
    }
    const int rc = factorize_dev(c, A, L);
    return rc < 0 ? 0 : 1;  // CHOLMOD returns TRUE even when the matrix is not positive definite
}

int nes_solve(int sys, nes_factor* L, const double* b, double* x, nes_ctx* c) {
    NES_ENTER(c);
    if (sys != 0) return fail(c, NES_ERR_INVALID, "nes_solve: only sys = 0 (CHOLMOD_A) is supported");
    if (!L || !b || !x) return fail(c, NES_ERR_INVALID, "nes_solve: null argument");
    NES_TRY(upload(c, L->d_rhs, b, L->m * sizeof(double)));
    const int rc = solve_dev(c, L, L->d_rhs);
    if (rc != 0) return rc;
    NES_TRY(download(c, x, L->d_rhs, L->m * sizeof(double)));
    return 0;
}

int nes_solve2(int sys, nes_factor* L, const double* b, double* x, nes_ctx* c) {
    // cholmod_solve2 differs from cholmod_solve only in reusing caller-held workspaces
    // (X, Y, E of solve-sparse-state); here they live inside the factor already.
    const int rc = nes_solve(sys, L, b, x, c);
    return rc == 0 ? 1 : 0;  // solve2 returns TRUE/FALSE (sparse-cholesky.lisp:546-554)
}

static void nes_free_factor_impl(nes_factor* L, nes_ctx* c) {
    if (!L->dense) sparse_free(c, L);
    dev_free(c, L->d_M);
    dev_free(c, L->d_dinv);
    dev_free(c, L->d_rhs);
    dev_free(c, L->d_info);
    dev_free(c, L->d_flags);
    dev_free(c, L->d_Winv);
    dev_free(c, L->d_As);
    dev_free(c, L->d_tile_list);
    if (L->dist) {
        dist_free_plan(c, L->dist);
        delete L->dist;
        L->dist = nullptr;
    }
    dev_free(c, L->d_defer_tiles);
    dev_free(c, L->d_defer_ws);
    dev_free(c, L->d_defer_counters);
}

int nes_free_factor(nes_factor** L, nes_ctx* c) {
    if (!c) return 0;
    if (!L || !*L) return 1;
    if (c->started) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
    }
    nes_free_factor_impl(*L, c);
    delete *L;
    *L = nullptr;
    return 1;
}

int nes_solve_dense(const double* B, size_t nrow, size_t ncol, const double* b, double* x,
                    nes_ctx* c) {
    NES_ENTER(c);
    nes_matrix* A = nes_dense_to_matrix(B, nrow, ncol, nrow, c);
    if (!A) return c->status;
    nes_factor* L = nes_analyze(A, c);
    int rc = L ? 0 : c->status;
    if (L) {
        rc = factorize_dev(c, A, L);
        if (rc == 0) rc = nes_solve(0, L, b, x, c);
    }
    nes_free_factor(&L, c);
    nes_free_matrix(&A, c);
    if (rc > 0) c->status = rc;  // NOT_POSDEF survives the frees, like cholmod_common.status
    return rc;
}

int nes_factor_to_dense(nes_factor* L, double* Lout, size_t ld, int* perm, nes_ctx* c) {
    NES_ENTER(c);
    if (!L || !Lout || ld < L->m) return fail(c, NES_ERR_INVALID, "nes_factor_to_dense: bad argument");
    if (!L->dense) return sparse_factor_to_dense(c, L, Lout, ld, perm);
    const size_t m = L->m;
    NES_CUDA(c, cudaMemcpy2DAsync(Lout, ld * sizeof(double), L->d_M, L->ld * sizeof(double),
                                  m * sizeof(double), m, cudaMemcpyDeviceToHost, c->stream));
    NES_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t j = 0; j < m; ++j)
        for (size_t i = 0; i < j; ++i) Lout[i + j * ld] = 0.0;
    if (perm)
        for (size_t i = 0; i < m; ++i) perm[i] = (int)i;
    return 0;
}

int nes_factor_residual(nes_matrix* A, nes_factor* L, double out[3], nes_ctx* c) {
    NES_ENTER(c);
    if (!A || !L || !out) return fail(c, NES_ERR_INVALID, "nes_factor_residual: null argument");
    if (!A->base->dense || !L->dense || A->base->m != L->m)
        return fail(c, NES_ERR_INVALID, "nes_factor_residual: dense matrix and its dense factor only");
    if (!L->factorized) return fail(c, NES_ERR_INVALID, "nes_factor_residual: factor is not factorized");
    const size_t m = L->m;
    const long long ld = (long long)L->ld;
    // T: a second factor object holds the freshly formed M (whole lower triangle on every rank: the check
    // runs replicated, every rank holds the complete factor after a distributed factorization)
    const int nranks = c->nranks;
    c->nranks = 1;
    nes_factor* T = nes_analyze(A, c);
    if (T && T->dist) {  // NES_FORCE_DIST: the check still forms the whole triangle with the plain enumeration
        dist_free_plan(c, T->dist);
        delete T->dist;
        T->dist = nullptr;
    }
    int rc = T ? dense_form_normal(c, A, T) : c->status;
    c->nranks = nranks;
    double* Lc = T ? static_cast<double*>(dev_alloc(c, (size_t)ld * m * sizeof(double))) : nullptr;
    const int G = c->num_sms * 4;
    double* part = T ? static_cast<double*>(dev_alloc(c, (size_t)G * sizeof(double))) : nullptr;
    std::vector<double> h(G);
    auto frob = [&](double* res) -> int {
        frob_lower_kernel<<<G, 256, 0, c->stream>>>(T->d_M, ld, (int)m, part);
        NES_CHECK_LAUNCH(c);
        NES_TRY(download(c, h.data(), part, (size_t)G * sizeof(double)));
        double v = 0.0;
        for (int i = 0; i < G; ++i) v += h[i];
        *res = sqrt(v);
        return 0;
    };
    double normM = 0.0, normR = 0.0;
    if (rc == 0 && (!Lc || !part)) rc = c->status;
    if (rc == 0) rc = frob(&normM);
    if (rc == 0) {
        tril_copy_kernel<<<G, 256, 0, c->stream>>>(L->d_M, Lc, ld, (int)m);
        ++c->launches;
        CUtensorMap mapL;
        if (make_operand_map(&mapL, Lc, (long long)m, (long long)m, ld) != 0)
            rc = fail(c, NES_ERR_CUDA, "cuTensorMapEncodeTiled failed for the copy of L");
        if (rc == 0) {  // T(lower tiles) -= Lc Lc'  on the FP64 tensor cores (K = m)
            NtArgs a{};
            a.C = T->d_M;
            a.ldc = ld;
            a.M = a.N = (int)m;
            a.K = (int)m;
            a.alpha = -1.0;
            a.beta = 1.0;
            a.lower = 1;
            a.same_operand = 1;
            cudaError_t e = nt_launch(mapL, mapL, a, c->num_sms, c->stream);
            ++c->launches;
            if (e != cudaSuccess) rc = fail(c, NES_ERR_CUDA, "residual product failed: %s", cudaGetErrorString(e));
        }
    }
    if (rc == 0) rc = frob(&normR);
    if (c->started) cudaStreamSynchronize(c->stream);
    dev_free(c, Lc);
    dev_free(c, part);
    nes_free_factor(&T, c);
    if (rc != 0) return rc;
    out[0] = normM > 0.0 ? normR / normM : normR;
    out[1] = normM;
    out[2] = normR;
    return 0;
}

int nes_normal_matrix_to_dense(nes_matrix* A, double* Mout, size_t ld, nes_ctx* c) {
    NES_ENTER(c);
    if (!A || !Mout || ld < A->base->m) return fail(c, NES_ERR_INVALID, "bad argument");
    if (!A->base->dense) return fail(c, NES_ERR_INVALID, "dense matrices only");
    nes_factor* L = nes_analyze(A, c);
    if (!L) return c->status;
    int rc = dense_form_normal(c, A, L);
    const size_t m = L->m;
    if (rc == 0) {
        cudaError_t e = cudaMemcpy2DAsync(Mout, ld * sizeof(double), L->d_M, L->ld * sizeof(double),
                                          m * sizeof(double), m, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(c, NES_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(e));
    }
    nes_free_factor(&L, c);
    return rc;
}

int nes_sdmult(nes_matrix* A, int transpose, const double alpha[2], const double beta[2],
               const double* x, double* y, nes_ctx* c) {
    if (!c || !c->started) return 0;
    cudaSetDevice(c->device);
    if (!A || !alpha || !beta || !x || !y) {
        fail(c, NES_ERR_INVALID, "nes_sdmult: null argument");
        return 0;
    }
    const size_t m = A->base->m, n = A->base->n;
    const size_t nx = transpose ? m : n, ny = transpose ? n : m;
    const size_t px = (nx + 1) / 2 * 2, py = (ny + 1) / 2 * 2;
    double* ws = ensure_ws(c, WS_API, (px + py) * sizeof(double));
    if (!ws) return 0;
    double* d_x = ws;
    double* d_y = ws + px;
    if (upload(c, d_x, x, nx * sizeof(double)) != 0) return 0;
    if (beta[0] != 0.0 && upload(c, d_y, y, ny * sizeof(double)) != 0) return 0;
    if (matvec(c, A, transpose, alpha[0], d_x, beta[0], d_y) != 0) return 0;
    if (download(c, y, d_y, ny * sizeof(double)) != 0) return 0;
    return 1;  // cholmod_sdmult returns TRUE on success (sparse-cholesky.lisp:600)
}

}  // extern "C"
