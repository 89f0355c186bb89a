// First-order solver of approx.lisp on the GPU: APPROX -- accelerated parallel
// proximal coordinate descent with full-vector steps -- on the penalised primal-dual formulation
// min sum_i 1/2 (scale_i (k_i . v - rhs_i))^2 + lin . v + complementarity terms,  l <= v <= u,
// v = [x | y | z | w].  The stacked constraint matrix K (one row per `quadratic`, approx.lisp:36-58) is an
// ordinary sparse nes_matrix, so value-&-gradient (approx.lisp:301-351) is two of the engine's
// atomic-free SpMV kernels (K v by rows, K' r by columns) plus elementwise / reduction kernels; every
// iteration of `approx` (:425-459) is HBM-bound: two gradients = 4 passes over K.
//   scale-quadratic (:70-74)        approx_scale_kernel
//   accumulate-nu (:97-113)         approx_nu_kernel
//   violation, %value-&-gradient    approx_resid_kernel + SpMVs + approx_comp_kernel
//   solve-coordinate, approx-iteration (:353-398)   approx_y_kernel, approx_descent_kernel
//   restart test, project-gradient (:400-423, :441-449)   approx_dot_kernel, approx_apply_kernel
// Reductions are two-stage (per-CTA partials, combined by one thread in block order): bitwise reproducible.
#include <cstdlib>

#include "nes_internal.h"

struct nes_approx {
    nes_matrix* K = nullptr;  // borrowed
    int R = 0, N = 0, ncomp = 0;
    double z0 = 0.0;
    double *rhs = nullptr, *scale = nullptr, *beta = nullptr, *lin = nullptr, *nu = nullptr, *l = nullptr, *u = nullptr;
    int *comp_x = nullptr, *comp_y = nullptr, *comp_flip = nullptr;
    double* comp_x0 = nullptr;
    // one dense quadratic kept out of K (the duality-gap row of make-approx touches every variable; a
    // thread-per-row SpMV would serialise on it): coefficients, rhs; dense[N] = scale, dense[N+1] = beta
    double* dense = nullptr;
    double dense_rhs = 0.0;
    int has_dense = 0;
    // iteration state
    double *x = nullptr, *z = nullptr, *y = nullptr, *zp = nullptr, *g = nullptr, *t = nullptr, *rs = nullptr;
    double* red = nullptr;  // 8 scalars
    double *partr = nullptr, *partn = nullptr;  // per-CTA partials of the row / variable reductions
    // variant 0 = approx.lisp; 1 = alm-approx.lisp: step damped by 0.95 (:198-216), stop when i > 10 and
    // |pg| < accuracy or at the last iteration (:328-333), the max ignores the linear term (:188-190)
    int variant = 0;
    double accuracy = 1e-10;
    cudaGraphExec_t graph = nullptr;  // one iteration of `approx`, captured per variant
    int graph_variant = -1;
    long long graph_launches = 0;
    bool graph_off = false;
};

namespace nes {


__global__ void approx_scale_kernel(int R, const int* __restrict__ rowptr, const double* __restrict__ val,
                                    const double* __restrict__ rhs, int do_scale, double* __restrict__ scale,
                                    double* __restrict__ beta) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double acc = 0.0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) acc = fma(val[k], val[k], acc);
    acc = fma(rhs[r], rhs[r], acc);
    const double norm = sqrt(acc);
    scale[r] = (do_scale && norm > 1e-6) ? 1.0 / norm : 1.0;
    beta[r] = (double)(rowptr[r + 1] - rowptr[r]);
}

__global__ void approx_nu_kernel(int N, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                                 const double* __restrict__ val, const double* __restrict__ scale,
                                 const double* __restrict__ beta, const double* __restrict__ dense,
                                 double* __restrict__ nu) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double acc = 0.0;
    if (dense) {
        const double cs = dense[j] * dense[N];
        acc = dense[N + 1] * cs * cs;
    }
    for (int k = colptr[j]; k < colptr[j + 1]; ++k) {
        const int r = rowidx[k];
        const double cs = val[k] * scale[r];
        acc = fma(beta[r], cs * cs, acc);
    }
    nu[j] = acc;
}

// scale-quadratic and beta of the dense row: dense[N] = scale, dense[N+1] = number of non-zero coefficients
__global__ void __launch_bounds__(1024)
approx_dense_scale_kernel(int N, double* __restrict__ dense, double rhs, int do_scale) {
    __shared__ double sh[1024];
    __shared__ double shc[1024];
    double a = 0.0, cnt = 0.0;
    for (int j = threadIdx.x; j < N; j += 1024) {
        a = fma(dense[j], dense[j], a);
        cnt += dense[j] != 0.0 ? 1.0 : 0.0;
    }
    sh[threadIdx.x] = a;
    shc[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[threadIdx.x] += sh[threadIdx.x + o];
            shc[threadIdx.x] += shc[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double norm = sqrt(sh[0] + rhs * rhs);
        dense[N] = (do_scale && norm > 1e-6) ? 1.0 / norm : 1.0;
        dense[N + 1] = shc[0];
    }
}

// ---- two-stage reductions: every CTA leaves up to 4 partial values in part[blockIdx.x][0..3], a single
// small CTA combines them in block order (fixed order => bitwise reproducible, no atomics)
constexpr int AP_B = 256;       // threads of the stage-1 kernels
constexpr int AP_MAXG = 1184;   // at most 8 CTAs per SM worth of partials

__device__ __forceinline__ double cta_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < AP_B / 32; ++w) r += sh[w];
    return r;  // valid in thread 0
}
__device__ __forceinline__ double cta_max(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < AP_B / 32; ++w) r = fmax(r, sh[w]);
    return r;
}

// rows: viol_r = (t_r - rhs_r) scale_r; rs_r = scale_r viol_r (factor of the gradient K' rs);
// partials: [0] sum 1/2 viol^2, [1] max 1/2 viol^2
__global__ void __launch_bounds__(AP_B)
approx_resid_kernel(int R, const double* __restrict__ t, const double* __restrict__ rhs,
                    const double* __restrict__ scale, double* __restrict__ rs, double* __restrict__ part) {
    __shared__ double sh[AP_B / 32];
    double sum = 0.0, mx = 0.0;
    for (int r = blockIdx.x * AP_B + threadIdx.x; r < R; r += gridDim.x * AP_B) {
        const double v = (t[r] - rhs[r]) * scale[r];
        rs[r] = scale[r] * v;
        const double val = 0.5 * v * v;
        sum += val;
        mx = fmax(mx, val);
    }
    const double s = cta_sum(sum, sh);
    const double m = cta_max(mx, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 4 + 0] = s;
        part[blockIdx.x * 4 + 1] = m;
    }
}

// variables: partials [0] lin . v, [1] dense . v, [2] sum of complementarity values, [3] their max
__global__ void __launch_bounds__(AP_B)
approx_ndot_kernel(int N, int ncomp, const double* __restrict__ lin, const double* __restrict__ dense,
                   const int* __restrict__ cx, const int* __restrict__ cy, const double* __restrict__ cx0,
                   const int* __restrict__ flip, const double* __restrict__ v, double* __restrict__ part) {
    __shared__ double sh[AP_B / 32];
    double ls = 0.0, ds = 0.0, cs = 0.0, cm = 0.0;
    for (int j = blockIdx.x * AP_B + threadIdx.x; j < N; j += gridDim.x * AP_B) {
        const double vj = v[j];
        ls = fma(lin[j], vj, ls);
        if (dense) ds = fma(dense[j], vj, ds);
    }
    for (int k = blockIdx.x * AP_B + threadIdx.x; k < ncomp; k += gridDim.x * AP_B) {
        double xk = v[cx[k]] - cx0[k];
        double yk = v[cy[k]];
        if (flip[k]) xk = -xk;
        xk = xk < 0.0 ? 0.0 : xk;
        yk = yk < 0.0 ? 0.0 : yk;
        const double val = yk * xk;
        cs += val;
        cm = fmax(cm, fabs(val));
    }
    const double a = cta_sum(ls, sh), b2 = cta_sum(ds, sh), c2 = cta_sum(cs, sh), d2 = cta_max(cm, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 4 + 0] = a;
        part[blockIdx.x * 4 + 1] = b2;
        part[blockIdx.x * 4 + 2] = c2;
        part[blockIdx.x * 4 + 3] = d2;
    }
}

// value-&-gradient scalars (approx.lisp:338-351): red[0] = sum of constraint values, red[1] = max |value|,
// red[5] = scale^2 * (dense . v - rhs) (multiplies the dense row in the gradient)
__global__ void __launch_bounds__(AP_B)
approx_value_finish_kernel(int gr, const double* __restrict__ partr, int gn, const double* __restrict__ partn,
                           const double* __restrict__ dense, int N, double dense_rhs, int max_skips_linear,
                           double* __restrict__ red) {
    __shared__ double sh[AP_B / 32];
    double qs = 0.0, qm = 0.0, ls = 0.0, ds = 0.0, cs = 0.0, cm = 0.0;
    for (int b = threadIdx.x; b < gr; b += AP_B) {
        qs += partr[b * 4 + 0];
        qm = fmax(qm, partr[b * 4 + 1]);
    }
    for (int b = threadIdx.x; b < gn; b += AP_B) {
        ls += partn[b * 4 + 0];
        ds += partn[b * 4 + 1];
        cs += partn[b * 4 + 2];
        cm = fmax(cm, partn[b * 4 + 3]);
    }
    qs = cta_sum(qs, sh);
    qm = cta_max(qm, sh);
    ls = cta_sum(ls, sh);
    ds = cta_sum(ds, sh);
    cs = cta_sum(cs, sh);
    cm = cta_max(cm, sh);
    if (threadIdx.x != 0) return;
    double tot = qs + cs + ls, mm = fmax(qm, cm), fac = 0.0;
    if (!max_skips_linear) mm = fmax(mm, fabs(ls));
    if (dense) {
        const double viol = (ds - dense_rhs) * dense[N];
        tot += 0.5 * viol * viol;
        mm = fmax(mm, 0.5 * viol * viol);
        fac = dense[N] * viol;
    }
    red[0] = tot;
    red[1] = mm;
    red[5] = fac;
    red[6] = ls;  // value of the linear constraint (dual-value of alm-approx.lisp:136-140 = z0 + this)
}

__global__ void approx_addlin_kernel(int N, const double* __restrict__ lin, const double* __restrict__ dense,
                                     const double* __restrict__ red, double* __restrict__ g) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N) g[j] += lin[j] + (dense ? dense[j] * red[5] : 0.0);
}

// theta lives in red[7] on the device so that one iteration is a fixed launch sequence (CUDA graph)
__global__ void approx_y_kernel(int N, const double* __restrict__ red, const double* __restrict__ x,
                                const double* __restrict__ z, double* __restrict__ y) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const double theta = red[7];
    if (j < N) y[j] = (1.0 - theta) * x[j] + theta * z[j];
}

// theta <- 1 on a restart, else 1/2 (sqrt(theta^4 + 4 theta^2) - theta^2) in the reference's operation
// order (approx.lisp:393-396 / alm-approx.lisp:258-263), without FMA contraction
__global__ void approx_theta_kernel(double* __restrict__ red, int variant) {
    const double t = red[7];
    const double t2 = __dmul_rn(t, t);
    const double rad = variant == 1 ? __dmul_rn(__dadd_rn(4.0, t2), t2)
                                    : __dadd_rn(__dmul_rn(t2, t2), __dmul_rn(4.0, t2));
    const double next = __dmul_rn(0.5, __dsub_rn(__dsqrt_rn(rad), t2));
    red[7] = red[2] > 0.0 ? 1.0 : next;
}

// zp = solve-coordinate (approx.lisp:353-369); x <- y + theta (zp - z) (approx-iteration :390-392)
__global__ void approx_descent_kernel(int N, const double* __restrict__ red, double damping,
                                      const double* __restrict__ y, const double* __restrict__ z,
                                      const double* __restrict__ nu, const double* __restrict__ g,
                                      const double* __restrict__ l, const double* __restrict__ u,
                                      double* __restrict__ zp, double* __restrict__ x) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const double theta = red[7];
    const double step = theta * nu[j];
    const double zj = z[j], gj = g[j];
    double best;
    if (step == 0.0) {
        best = gj < 0.0 ? u[j] : (gj > 0.0 ? l[j] : zj);
    } else {
        best = zj - damping * (gj / step);
        if (best < l[j]) best = l[j];
        else if (best > u[j]) best = u[j];
    }
    zp[j] = best;
    x[j] = y[j] + theta * (best - zj);
}

// partials [0] g . (zp - z) (dot-diff, :412-417), [1] g . g
__global__ void __launch_bounds__(AP_B)
approx_dot_kernel(int N, const double* __restrict__ g, const double* __restrict__ z, const double* __restrict__ zp,
                  double* __restrict__ part) {
    __shared__ double sh[AP_B / 32];
    double d = 0.0, gg = 0.0;
    for (int j = blockIdx.x * AP_B + threadIdx.x; j < N; j += gridDim.x * AP_B) {
        d = fma(g[j], zp[j] - z[j], d);
        gg = fma(g[j], g[j], gg);
    }
    const double a = cta_sum(d, sh), b2 = cta_sum(gg, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 4 + 0] = a;
        part[blockIdx.x * 4 + 1] = b2;
    }
}

// restart (x <- z) or accept (z <- zp) (:441-446); partial [0] of |z - clamp(z - g)|^2 (project-gradient)
__global__ void __launch_bounds__(AP_B)
approx_apply_kernel(int N, const double* __restrict__ red, const double* __restrict__ g,
                    const double* __restrict__ zp, const double* __restrict__ l, const double* __restrict__ u,
                    double* __restrict__ z, double* __restrict__ x, double* __restrict__ part) {
    __shared__ double sh[AP_B / 32];
    const bool restart = red[2] > 0.0;  // dot-diff of this iteration, decided on the device: one host sync less
    double acc = 0.0;
    for (int j = blockIdx.x * AP_B + threadIdx.x; j < N; j += gridDim.x * AP_B) {
        double zj;
        if (restart) {
            zj = z[j];
            x[j] = zj;
        } else {
            zj = zp[j];
            z[j] = zj;
        }
        double xp = zj - g[j];
        if (xp < l[j]) xp = l[j];
        if (xp > u[j]) xp = u[j];
        const double dd = zj - xp;
        acc = fma(dd, dd, acc);
    }
    const double a = cta_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x * 4 + 0] = a;
}

// red[o0] = sum of partial 0 (sqrt if sq0), red[o1] = sqrt(sum of partial 1) when o1 >= 0 (one CTA, fixed tree)
__global__ void __launch_bounds__(AP_B)
approx_finish_kernel(int g, const double* __restrict__ part, int o0, int sq0, int o1, double* __restrict__ red) {
    __shared__ double sh[AP_B / 32];
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < g; k += AP_B) {
        a += part[k * 4 + 0];
        b += part[k * 4 + 1];
    }
    a = cta_sum(a, sh);
    b = cta_sum(b, sh);
    if (threadIdx.x != 0) return;
    red[o0] = sq0 ? sqrt(a) : a;
    if (o1 >= 0) red[o1] = sqrt(b);
}

// complementarity constraints of one orientation (each x_i and each y_i appears at most once per
// orientation, so plain read-modify-writes do not collide)
__global__ void approx_comp_kernel(int ncomp, const int* __restrict__ cx, const int* __restrict__ cy,
                                   const double* __restrict__ cx0, const int* __restrict__ flip, int want_flip,
                                   const double* __restrict__ v, double* __restrict__ g) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncomp || flip[k] != want_flip) return;
    double xk = v[cx[k]] - cx0[k];
    double yk = v[cy[k]];
    if (want_flip) xk = -xk;
    xk = xk < 0.0 ? 0.0 : xk;
    yk = yk < 0.0 ? 0.0 : yk;
    g[cx[k]] += want_flip ? -yk : yk;
    g[cy[k]] += xk;
}

__global__ void approx_fill_kernel(int n, double v, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

__global__ void approx_project_kernel(int N, const double* __restrict__ l, const double* __restrict__ u,
                                      double* __restrict__ x, double* __restrict__ z) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    double v = x[j];
    v = fmin(u[j], fmax(l[j], v));
    x[j] = v;
    z[j] = v;
}

static int ap_grid(nes_ctx* c, int n) {
    int g = (n + AP_B - 1) / AP_B;
    const int cap = c->num_sms * 4 < AP_MAXG ? c->num_sms * 4 : AP_MAXG;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

// g <- gradient at v, red[0..1] <- value, max
static int approx_value_gradient_dev(nes_ctx* c, nes_approx* st, const double* d_v) {
    const MatrixBase* b = st->K->base;
    const double* dense = st->has_dense ? st->dense : nullptr;
    NES_TRY(matvec_unscaled(c, b, 0, 1.0, d_v, 0.0, st->t));
    {
        StageTimer timer(c, NES_STAGE_VECTOR);
        const int gr = ap_grid(c, st->R), gn = ap_grid(c, st->N);
        approx_resid_kernel<<<gr, AP_B, 0, c->stream>>>(st->R, st->t, st->rhs, st->scale, st->rs, st->partr);
        NES_CHECK_LAUNCH(c);
        approx_ndot_kernel<<<gn, AP_B, 0, c->stream>>>(st->N, st->ncomp, st->lin, dense, st->comp_x, st->comp_y,
                                                      st->comp_x0, st->comp_flip, d_v, st->partn);
        NES_CHECK_LAUNCH(c);
        approx_value_finish_kernel<<<1, AP_B, 0, c->stream>>>(gr, st->partr, gn, st->partn, dense, st->N, st->dense_rhs,
                                                             st->variant == 1, st->red);
        NES_CHECK_LAUNCH(c);
    }
    NES_TRY(matvec_unscaled(c, b, 1, 1.0, st->rs, 0.0, st->g));
    StageTimer timer(c, NES_STAGE_VECTOR);
    approx_addlin_kernel<<<(st->N + 255) / 256, 256, 0, c->stream>>>(st->N, st->lin, dense, st->red, st->g);
    NES_CHECK_LAUNCH(c);
    if (st->ncomp > 0) {
        for (int flip = 0; flip < 2; ++flip) {
            approx_comp_kernel<<<(st->ncomp + 255) / 256, 256, 0, c->stream>>>(
                st->ncomp, st->comp_x, st->comp_y, st->comp_x0, st->comp_flip, flip, d_v, st->g);
            NES_CHECK_LAUNCH(c);
        }
    }
    return 0;
}

template <typename T>
static T* ap_upload(nes_ctx* c, const T* h, size_t n) {
    T* p = static_cast<T*>(dev_alloc(c, (n + 2) * sizeof(T)));
    if (!p) return nullptr;
    if (n && h && upload(c, p, h, n * sizeof(T)) != 0) return nullptr;
    return p;
}

}  // namespace nes

using namespace nes;

extern "C" {

nes_approx* nes_approx_create(nes_matrix* K, const double* rhs, const double* lin, const double* l, const double* u,
                              const int* comp_x, const int* comp_y, const double* comp_x0, const int* comp_flipped,
                              int ncomp, const double* dense_row, double dense_rhs, int scale, double z0,
                              nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (!K || !K->base || K->base->dense || !rhs || !lin || !l || !u || ncomp < 0) {
        fail(c, NES_ERR_INVALID, "nes_approx_create: needs a sparse constraint matrix and its vectors");
        return nullptr;
    }
    nes_approx* st = new nes_approx();
    st->K = K;
    st->R = (int)K->base->m;
    st->N = (int)K->base->n;
    st->ncomp = ncomp;
    st->z0 = z0;
    const size_t R = st->R, N = st->N;
    st->rhs = ap_upload(c, rhs, R);
    st->lin = ap_upload(c, lin, N);
    st->l = ap_upload(c, l, N);
    st->u = ap_upload(c, u, N);
    st->comp_x = ap_upload(c, comp_x, (size_t)ncomp);
    st->comp_y = ap_upload(c, comp_y, (size_t)ncomp);
    st->comp_x0 = ap_upload(c, comp_x0, (size_t)ncomp);
    st->comp_flip = ap_upload(c, comp_flipped, (size_t)ncomp);
    st->has_dense = dense_row != nullptr;
    st->dense_rhs = dense_rhs;
    st->dense = ap_upload(c, dense_row, N);  // two extra slots hold its scale and beta
    st->scale = ap_upload<double>(c, nullptr, R);
    st->beta = ap_upload<double>(c, nullptr, R);
    st->nu = ap_upload<double>(c, nullptr, N);
    st->x = ap_upload<double>(c, nullptr, N);
    st->z = ap_upload<double>(c, nullptr, N);
    st->y = ap_upload<double>(c, nullptr, N);
    st->zp = ap_upload<double>(c, nullptr, N);
    st->g = ap_upload<double>(c, nullptr, N);
    st->t = ap_upload<double>(c, nullptr, R);
    st->rs = ap_upload<double>(c, nullptr, R);
    st->red = ap_upload<double>(c, nullptr, 8);
    st->partr = ap_upload<double>(c, nullptr, (size_t)AP_MAXG * 4);
    st->partn = ap_upload<double>(c, nullptr, (size_t)AP_MAXG * 4);
    void* all[] = {st->rhs, st->lin, st->l, st->u, st->comp_x, st->comp_y, st->comp_x0, st->comp_flip, st->scale,
                   st->beta, st->nu, st->x, st->z, st->y, st->zp, st->g, st->t, st->rs, st->red, st->dense, st->partr,
                   st->partn};
    for (void* p : all)
        if (!p) {
            nes_approx* tmp = st;
            nes_approx_free(&tmp, c);
            return nullptr;
        }
    const MatrixBase* b = K->base;
    approx_scale_kernel<<<(st->R + 255) / 256, 256, 0, c->stream>>>(st->R, b->d_rowptr, b->d_csr_val, st->rhs, scale,
                                                                   st->scale, st->beta);
    ++c->launches;
    if (st->has_dense) {
        approx_dense_scale_kernel<<<1, 1024, 0, c->stream>>>(st->N, st->dense, dense_rhs, scale);
        ++c->launches;
    }
    approx_nu_kernel<<<(st->N + 255) / 256, 256, 0, c->stream>>>(st->N, b->d_colptr, b->d_rowidx, b->d_values,
                                                                st->scale, st->beta,
                                                                st->has_dense ? st->dense : nullptr, st->nu);
    ++c->launches;
    if (cudaGetLastError() != cudaSuccess) {
        fail(c, NES_ERR_CUDA, "nes_approx_create: kernel launch failed");
        nes_approx* tmp = st;
        nes_approx_free(&tmp, c);
        return nullptr;
    }
    return st;
}

int nes_approx_free(nes_approx** pst, nes_ctx* c) {
    if (!c) return 0;
    if (!pst || !*pst) return 1;
    nes_approx* st = *pst;
    if (c->started) cudaSetDevice(c->device);
    void* all[] = {st->rhs, st->lin, st->l, st->u, st->comp_x, st->comp_y, st->comp_x0, st->comp_flip, st->scale,
                   st->beta, st->nu, st->x, st->z, st->y, st->zp, st->g, st->t, st->rs, st->red, st->dense, st->partr,
                   st->partn};
    for (void* p : all) dev_free(c, p);
    if (st->graph) cudaGraphExecDestroy(st->graph);
    delete st;
    *pst = nullptr;
    return 1;
}

int nes_approx_value_gradient(nes_approx* st, const double* x, double* value, double* g, double* maxv, nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !x) return fail(c, NES_ERR_INVALID, "nes_approx_value_gradient: null argument");
    NES_TRY(upload(c, st->y, x, (size_t)st->N * sizeof(double)));
    NES_TRY(approx_value_gradient_dev(c, st, st->y));
    double red[2];
    NES_TRY(download(c, red, st->red, sizeof(red)));
    if (value) *value = red[0];
    if (maxv) *maxv = red[1];
    if (g) NES_TRY(download(c, g, st->g, (size_t)st->N * sizeof(double)));
    return 0;
}

int nes_approx_get(nes_approx* st, int which, double* out, nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_approx_get: null argument");
    switch (which) {
        case 'n': return download(c, out, st->nu, (size_t)st->N * sizeof(double));
        case 's': return download(c, out, st->scale, (size_t)st->R * sizeof(double));
        case 'd': return st->has_dense ? download(c, out, st->dense + st->N, 2 * sizeof(double))
                                       : fail(c, NES_ERR_INVALID, "nes_approx_get: no dense row");
        case 'z': return download(c, out, st->z, (size_t)st->N * sizeof(double));
        case 'x': return download(c, out, st->x, (size_t)st->N * sizeof(double));
        default: return fail(c, NES_ERR_INVALID, "nes_approx_get: bad selector");
    }
}

// approx (approx.lisp:425-459, alm-approx.lisp:307-346).  stats[7] = {|g|_2, projected-gradient norm, max
// constraint value, value + z0, last dot-diff, theta, z0 + linear value} of the last iteration.
int nes_approx_solve(nes_approx* st, int n_iter, const double* x0, double* z_out, int* iters, int* restarts,
                     double* stats, nes_ctx* c) {
    NES_ENTER(c);
    if (!st) return fail(c, NES_ERR_INVALID, "nes_approx_solve: null state");
    const int N = st->N;
    const int nb = (N + 255) / 256;
    if (x0) NES_TRY(upload(c, st->x, x0, (size_t)N * sizeof(double)));
    else NES_CUDA(c, cudaMemsetAsync(st->x, 0, (size_t)N * sizeof(double), c->stream));
    approx_project_kernel<<<nb, 256, 0, c->stream>>>(N, st->l, st->u, st->x, st->z);
    NES_CHECK_LAUNCH(c);
    approx_fill_kernel<<<1, 32, 0, c->stream>>>(1, 1.0, st->red + 7);  // theta = 1
    NES_CHECK_LAUNCH(c);
    int nrestart = 0, done_at = n_iter;
    double red[8] = {0};
    double lin_value = 0.0, theta = 1.0;
    // one iteration = a fixed sequence of 19 launches: captured once, replayed (every iteration ends in a
    // host synchronisation, so the enqueue cost would otherwise sit on the critical path)
    auto enqueue_iteration = [&]() -> int {
        approx_y_kernel<<<nb, 256, 0, c->stream>>>(N, st->red, st->x, st->z, st->y);
        NES_CHECK_LAUNCH(c);
        NES_TRY(approx_value_gradient_dev(c, st, st->y));
        approx_descent_kernel<<<nb, 256, 0, c->stream>>>(N, st->red, st->variant == 1 ? 0.95 : 1.0, st->y, st->z,
                                                        st->nu, st->g, st->l, st->u, st->zp, st->x);
        NES_CHECK_LAUNCH(c);
        NES_TRY(approx_value_gradient_dev(c, st, st->zp));
        const int gn = ap_grid(c, N);
        approx_dot_kernel<<<gn, AP_B, 0, c->stream>>>(N, st->g, st->z, st->zp, st->partn);
        NES_CHECK_LAUNCH(c);
        approx_finish_kernel<<<1, AP_B, 0, c->stream>>>(gn, st->partn, 2, 0, 3, st->red);
        NES_CHECK_LAUNCH(c);
        approx_apply_kernel<<<gn, AP_B, 0, c->stream>>>(N, st->red, st->g, st->zp, st->l, st->u, st->z, st->x,
                                                        st->partn);
        NES_CHECK_LAUNCH(c);
        approx_finish_kernel<<<1, AP_B, 0, c->stream>>>(gn, st->partn, 4, 1, -1, st->red);
        NES_CHECK_LAUNCH(c);
        approx_theta_kernel<<<1, 1, 0, c->stream>>>(st->red, st->variant);
        NES_CHECK_LAUNCH(c);
        return 0;
    };
    static const bool env_off = getenv("NES_NO_GRAPH") != nullptr;
    const bool want_graph = !env_off && !st->graph_off && !c->timing;
    if (want_graph && st->graph && st->graph_variant != st->variant) {
        cudaGraphExecDestroy(st->graph);
        st->graph = nullptr;
    }
    if (want_graph && !st->graph) {
        const long long l0 = c->launches;
        cudaGraph_t graph = nullptr;
        int rc = -1;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = enqueue_iteration();
            cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            if (rc == 0 && e == cudaSuccess && graph && cudaGraphInstantiate(&st->graph, graph, 0) != cudaSuccess)
                st->graph = nullptr;
            if (graph) cudaGraphDestroy(graph);
        }
        st->graph_launches = c->launches - l0;
        c->launches = l0;
        if (!st->graph) {
            cudaGetLastError();
            st->graph_off = true;
            c->status = 0;
        } else {
            st->graph_variant = st->variant;
        }
    }
    for (int i = 0; i < n_iter; ++i) {
        if (want_graph && st->graph) {
            NES_CUDA(c, cudaGraphLaunch(st->graph, c->stream));
            c->launches += st->graph_launches;
        } else {
            NES_TRY(enqueue_iteration());
        }
        NES_TRY(download(c, red, st->red, 8 * sizeof(double)));  // the iteration's only host sync
        lin_value = red[6];
        theta = red[7];
        if (red[2] > 0.0) ++nrestart;
        const bool done = st->variant == 1 ? ((i > 10 && red[4] < st->accuracy) || i == n_iter - 1)
                                           : red[4] < 1e-10;
        if (done) {
            done_at = i + 1;
            break;
        }
    }
    if (iters) *iters = done_at;
    if (restarts) *restarts = nrestart;
    if (stats) {
        stats[0] = red[3];
        stats[1] = red[4];
        stats[2] = red[1];
        stats[3] = red[0] + st->z0;
        stats[4] = red[2];
        stats[5] = theta;
        stats[6] = st->z0 + lin_value;  // dual-value at the last zp (alm-approx.lisp:136-140, :340)
    }
    if (z_out) NES_TRY(download(c, z_out, st->z, (size_t)N * sizeof(double)));
    return 0;
}

int nes_approx_set_variant(nes_approx* st, int variant, double accuracy, nes_ctx* c) {
    NES_ENTER(c);
    if (!st || variant < 0 || variant > 1) return fail(c, NES_ERR_INVALID, "nes_approx_set_variant: bad argument");
    st->variant = variant;
    st->accuracy = accuracy;
    return 0;
}

// make-alm-subproblem (alm-approx.lisp:355-403) over the same constraint rows: every quadratic gets the
// scale sqrt(weight), the linear term becomes c + A' lambda (computed by the caller with nes_sdmult),
// z0 = -lambda . b; nu is recomputed (accumulate-nu).
int nes_approx_set_subproblem(nes_approx* st, double row_scale, const double* lin, double z0, nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !lin) return fail(c, NES_ERR_INVALID, "nes_approx_set_subproblem: null argument");
    const MatrixBase* b = st->K->base;
    NES_TRY(upload(c, st->lin, lin, (size_t)st->N * sizeof(double)));
    st->z0 = z0;
    approx_fill_kernel<<<(st->R + 255) / 256, 256, 0, c->stream>>>(st->R, row_scale, st->scale);
    NES_CHECK_LAUNCH(c);
    approx_nu_kernel<<<(st->N + 255) / 256, 256, 0, c->stream>>>(st->N, b->d_colptr, b->d_rowidx, b->d_values,
                                                                st->scale, st->beta,
                                                                st->has_dense ? st->dense : nullptr, st->nu);
    NES_CHECK_LAUNCH(c);
    return 0;
}

}  // extern "C"
