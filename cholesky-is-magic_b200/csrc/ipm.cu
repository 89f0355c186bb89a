// Interior-point layer on the device: solve-kkt-newton and the primal-dual affine scaling state.
//
//   kkt_newton_dev          newton-solve.lisp:139-154 / sparse-newton-solve.lisp:150-168
//   nes_pdas_violation      primal-dual-affine-scaling.lisp:135-150 (+ scalars of :325-332)
//   nes_pdas_newton_direction  :152-164 direction, :166-198 box-step / pos-step / pdas-step
//   nes_pdas_apply_step     :200-207
//   nes_pdas_repair         :268-288 one-repair-iteration (cholesky-ls! :223-233, slack :235-246,
//                           max-step :253-266)
//   nes_pdas_recentre       :348-366 (centering-direction :290-303, primal-project :305-317)
//   nes_pdas_one_iteration  :319-383,  nes_pdas_solve  :385-396
//
// All vectors stay on the GPU; each call returns a handful of scalars through one pinned copy.
#include <cmath>
#include <limits>

#include "ipm_kernels.cuh"
#include "nes_internal.h"

using namespace nes;

namespace nes {
int factorize_dev(nes_ctx* c, nes_matrix* A, nes_factor* L);  // nes_factor.cu
int solve_dev(nes_ctx* c, nes_factor* L, double* d_x);

static inline int vec_grid(const nes_ctx* c, size_t n) {
    size_t g = (n + RED_THREADS - 1) / RED_THREADS;
    size_t cap = (size_t)c->num_sms * 2;
    if (g > cap) g = cap;
    return g ? (int)g : 1;
}

// ---------------------------------------------------------------------------------------------
// solve-kkt-newton, elementwise parts
// ---------------------------------------------------------------------------------------------
// newton-solve.lisp:27-33, 58-59, 64-66, 92-98 in one pass (inputs untouched):
//   [filter-U/Z]  w'=w/u  e'=e/u  l'=l/z  f'=f/z  h1=h+e'  q=1+w'l'  h2=h1+w'f'  d=1/q  h3=d h2
//   s=sqrt(l' d)  v = l' h3 - f'   (so that g2 = g + A v fuses the two forward products)
__global__ void kkt_pre_kernel(size_t n, int filters, const double* __restrict__ l_in,
                               const double* __restrict__ u_in, const double* __restrict__ w_in,
                               const double* __restrict__ z_in, const double* __restrict__ e_in,
                               const double* __restrict__ f_in, const double* __restrict__ h_in,
                               double* __restrict__ wp, double* __restrict__ ep, double* __restrict__ lp,
                               double* __restrict__ fp, double* __restrict__ d_out,
                               double* __restrict__ h3, double* __restrict__ s_out,
                               double* __restrict__ v_out, int* __restrict__ div0) {
    bool zero_div = false;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        double l = l_in[i], u = u_in[i], w = w_in[i], z = z_in[i], e = e_in[i], f = f_in[i], h = h_in[i];
        if (filters) {
            if (u > 1e7) {  // filter-U, sparse-newton-solve.lisp:30-38
                u = 1.0;
                e = w;
                w = 0.0;
            }
            if (l > 1e7) {  // filter-Z, sparse-newton-solve.lisp:40-45
                l = 1.0;
                f = z;
                z = 0.0;
            }
        }
        // scale-U / scale-Z divide by u and z: filter-Z leaves z = 0 behind (the reference traps here with
        // DIVISION-BY-ZERO under SBCL); the flag turns the resulting NaN factorization into an explicit status
        zero_div |= (u == 0.0) | (z == 0.0);
        const double iu = 1.0 / u;
        w = iu * w;
        e = iu * e;
        const double iz = 1.0 / z;
        l = iz * l;
        f = iz * f;
        h = e + h;
        const double q = 1.0 + w * l;
        h = w * f + h;
        const double d = 1.0 / q;
        h = d * h;
        wp[i] = w;
        ep[i] = e;
        lp[i] = l;
        fp[i] = f;
        d_out[i] = d;
        h3[i] = h;
        s_out[i] = sqrt(l * d);
        v_out[i] = l * h - f;
    }
    if (zero_div) *div0 = 1;
}

// newton-solve.lisp:127-137: dz = h3 - d (A'dy);  dx = f' - dz l';  dw = dx w' + e'.
// Optionally fused with pdas-step (primal-dual-affine-scaling.lisp:166-198): block partial of
// min( l/dx [dx>0], u/-dx [dx<0], w/dw [dw>0], z/dz [dz>0] ).
__global__ void kkt_post_kernel(size_t n, const double* __restrict__ r, const double* __restrict__ h3,
                                const double* __restrict__ d, const double* __restrict__ fp,
                                const double* __restrict__ lp, const double* __restrict__ wp,
                                const double* __restrict__ ep, double* __restrict__ dw,
                                double* __restrict__ dx, double* __restrict__ dz,
                                const double* __restrict__ l0, const double* __restrict__ u0,
                                const double* __restrict__ w0, const double* __restrict__ z0,
                                double* __restrict__ partial) {
    __shared__ double buf[32];
    double mn = INFINITY;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double vz = h3[i] - d[i] * r[i];
        const double vx = fp[i] - vz * lp[i];
        const double vw = vx * wp[i] + ep[i];
        dz[i] = vz;
        dx[i] = vx;
        dw[i] = vw;
        if (partial) {
            if (vx > 0.0) mn = fmin(mn, l0[i] / vx);
            else if (vx < 0.0) mn = fmin(mn, u0[i] / (-vx));
            if (vw > 0.0) mn = fmin(mn, w0[i] / vw);
            if (vz > 0.0) mn = fmin(mn, z0[i] / vz);
        }
    }
    if (partial) {
        mn = block_reduce(mn, RED_MIN, buf);
        if (threadIdx.x == 0) partial[blockIdx.x] = mn;
    }
}

struct KktVectors {
    // inputs (n, n, n, n, n, n, n) + g (m, overwritten with g2 then left alone)
    const double *l, *u, *w, *z, *e, *f, *h;
    double* g;
    // temporaries (n)
    double *wp, *ep, *lp, *fp, *d, *h3, *s, *v, *r;
    // outputs
    double *dw, *dx, *dz;  // n
    double* dy;            // m
    int* div0;             // device flag: scale-U / scale-Z divided by zero (filter-Z, see kkt_pre_kernel)
};

// Returns 0, NES_NOT_POSDEF, or a negative error.  `step_partial` (may be null) enables the fused
// pdas-step partial minimum using the original l0,u0,w0,z0 = kv.l, kv.u, kv.w, kv.z.
static int kkt_newton_dev(nes_ctx* c, nes_matrix* A, nes_factor* L, int filters, const KktVectors& kv,
                          double* step_partial, int* step_blocks) {
    const size_t n = A->base->n, m = A->base->m;
    const int grid = vec_grid(c, n);
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        NES_CUDA(c, cudaMemsetAsync(kv.div0, 0, sizeof(int), c->stream));
        kkt_pre_kernel<<<grid, RED_THREADS, 0, c->stream>>>(n, filters, kv.l, kv.u, kv.w, kv.z, kv.e,
                                                            kv.f, kv.h, kv.wp, kv.ep, kv.lp, kv.fp, kv.d,
                                                            kv.h3, kv.s, kv.v, kv.div0);
        NES_CHECK_LAUNCH(c);
    }
    // solve-delta-y: scale the columns of (a copy of) A by s, factorize (A s)(A s)', solve
    NES_TRY(set_scale_dev(c, A, kv.s));
    // g2 = g - A f' + A (l' h3) = g + A v      (newton-solve.lisp:61-62, 96-98)
    NES_TRY(matvec_unscaled(c, A->base, 0, 1.0, kv.v, 1.0, kv.g));
    const int rc = factorize_dev(c, A, L);
    if (rc == NES_NOT_POSDEF) {
        // a zero divisor made theta NaN, so the first pivot failed: report the cause, not " singular "
        int flag = 0;
        NES_TRY(download(c, &flag, kv.div0, sizeof(int)));
        if (flag) {
            fail(c, NES_DIV_BY_ZERO, "solve-kkt-newton: division by zero in scale-U/scale-Z (filter-Z sets z = 0 "
                                     "for x - lo > 1e7, sparse-newton-solve.lisp:40-45, then scale-Z divides by it)");
            return NES_DIV_BY_ZERO;
        }
    }
    if (rc != 0) return rc;
    NES_CUDA(c, cudaMemcpyAsync(kv.dy, kv.g, m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    NES_TRY(solve_dev(c, L, kv.dy));
    // solve-delta-z needs A' dy
    NES_TRY(matvec_unscaled(c, A->base, 1, 1.0, kv.dy, 0.0, kv.r));
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        kkt_post_kernel<<<grid, RED_THREADS, 0, c->stream>>>(n, kv.r, kv.h3, kv.d, kv.fp, kv.lp, kv.wp,
                                                             kv.ep, kv.dw, kv.dx, kv.dz, kv.l, kv.u,
                                                             kv.w, kv.z, step_partial);
        NES_CHECK_LAUNCH(c);
    }
    if (step_blocks) *step_blocks = grid;
    return 0;
}

static inline size_t pad2(size_t n) { return (n + 15) / 16 * 16; }

}  // namespace nes

// ---------------------------------------------------------------------------------------------
// PDAS state
// ---------------------------------------------------------------------------------------------
struct nes_pdas {
    nes_matrix* A = nullptr;  // own shallow copy (shares values, own column scale)
    nes_factor* L = nullptr;
    size_t n = 0, m = 0;
    int filters = 0;
    double* d_block = nullptr;  // one allocation for all vectors
    // n-vectors
    double *c, *lo, *hi, *x, *w, *z;
    double *l, *u, *wu, *zl, *rd;
    double *wp, *ep, *lp, *fp, *d, *h3, *s, *v, *r;
    double *dw, *dx, *dz;
    double *slack, *gdir, *tn;
    // m-vectors
    double *b, *y, *rp, *dy, *tm;
    // reduction scratch + device scalars
    double* partial;
    double* scal;  // 16 doubles
    int have_direction = 0;
};

namespace nes {

// violation, n-part.  partials: 0 c.x  1 lo.z  2 hi.w  3 max|rd|  4 max|wu|  5 max|zl|  6 min l  7 min u
__global__ void pdas_violation_n_kernel(size_t n, const double* __restrict__ x,
                                        const double* __restrict__ lo, const double* __restrict__ hi,
                                        const double* __restrict__ w, const double* __restrict__ z,
                                        const double* __restrict__ cvec, const double* __restrict__ aty,
                                        double* __restrict__ l, double* __restrict__ u,
                                        double* __restrict__ wu, double* __restrict__ zl,
                                        double* __restrict__ rd, double* __restrict__ partial) {
    __shared__ double buf[32];
    double acc[8] = {0.0, 0.0, 0.0, -INFINITY, -INFINITY, -INFINITY, INFINITY, INFINITY};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double xi = x[i], li = xi - lo[i], ui = hi[i] - xi, wi = w[i], zi = z[i];
        const double wui = wi * ui, zli = zi * li;
        const double rdi = (zi + aty[i]) - (wi + cvec[i]);
        l[i] = li;
        u[i] = ui;
        wu[i] = wui;
        zl[i] = zli;
        rd[i] = rdi;
        acc[0] = fma(cvec[i], xi, acc[0]);
        acc[1] = fma(lo[i], zi, acc[1]);
        acc[2] = fma(hi[i], wi, acc[2]);
        acc[3] = fmax(acc[3], fabs(rdi));
        acc[4] = fmax(acc[4], fabs(wui));
        acc[5] = fmax(acc[5], fabs(zli));
        acc[6] = fmin(acc[6], li);
        acc[7] = fmin(acc[7], ui);
    }
    const int ops[8] = {RED_SUM, RED_SUM, RED_SUM, RED_MAX, RED_MAX, RED_MAX, RED_MIN, RED_MIN};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double v = block_reduce(acc[k], ops[k], buf);
        if (threadIdx.x == 0) partial[k * gridDim.x + blockIdx.x] = v;
    }
}

// violation, m-part: rp = ax - b (in place on ax).  partials: 0 b.y  1 max|rp|
__global__ void pdas_violation_m_kernel(size_t m, double* __restrict__ ax_rp, const double* __restrict__ b,
                                        const double* __restrict__ y, double* __restrict__ partial) {
    __shared__ double buf[32];
    double by = 0.0, mx = -INFINITY;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m;
         i += (size_t)gridDim.x * blockDim.x) {
        const double rp = ax_rp[i] - b[i];
        ax_rp[i] = rp;
        by = fma(b[i], y[i], by);
        mx = fmax(mx, fabs(rp));
    }
    double v = block_reduce(by, RED_SUM, buf);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
    v = block_reduce(mx, RED_MAX, buf);
    if (threadIdx.x == 0) partial[gridDim.x + blockIdx.x] = v;
}

// out[0]=pobj out[1]=dobj out[2..5]=violations out[6]=min l out[7]=min u
__global__ void pdas_violation_finish_kernel(const double* __restrict__ pn, int gn,
                                             const double* __restrict__ pm, int gm,
                                             double* __restrict__ out) {
    if (threadIdx.x != 0) return;
    double r[8];
    const int ops[8] = {RED_SUM, RED_SUM, RED_SUM, RED_MAX, RED_MAX, RED_MAX, RED_MIN, RED_MIN};
    for (int k = 0; k < 8; ++k) {
        double v = red_identity(ops[k]);
        for (int b = 0; b < gn; ++b) v = red_combine(v, pn[k * gn + b], ops[k]);
        r[k] = v;
    }
    double by = 0.0, mrp = -INFINITY;
    for (int b = 0; b < gm; ++b) {
        by += pm[b];
        mrp = fmax(mrp, pm[gm + b]);
    }
    out[0] = r[0];
    // dobj = b.y + lo.z - hi.w  (primal-dual-affine-scaling.lisp:326-328)
    out[1] = (by + r[1]) + (-r[2]);
    out[2] = mrp;
    out[3] = r[3];
    out[4] = r[4];
    out[5] = r[5];
    out[6] = r[6];
    out[7] = r[7];
}

__global__ void axpy_neg_kernel(size_t n, double alpha, const double* __restrict__ dv,
                                double* __restrict__ v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        v[i] = fma(-alpha, dv[i], v[i]);
}

// apply-step: w,x,z (n) and y (m) in one launch
__global__ void apply_step_kernel(size_t n, size_t m, double alpha, const double* __restrict__ dw,
                                  const double* __restrict__ dx, const double* __restrict__ dz,
                                  const double* __restrict__ dy, double* __restrict__ w,
                                  double* __restrict__ x, double* __restrict__ z,
                                  double* __restrict__ y) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        w[i] = fma(-alpha, dw[i], w[i]);
        x[i] = fma(-alpha, dx[i], x[i]);
        z[i] = fma(-alpha, dz[i], z[i]);
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m; i += stride)
        y[i] = fma(-alpha, dy[i], y[i]);
}

// slack (primal-dual-affine-scaling.lisp:235-246): min(cap, x-lo, hi-x); partial 0 = min slack
// optionally also bumps w,z by `bump` (recentring branch, :349-350) and writes the centering
// direction (:290-303) scaled: sc = slack * (-cdir).
__global__ void slack_kernel(size_t n, double cap, const double* __restrict__ x,
                             const double* __restrict__ lo, const double* __restrict__ hi,
                             double* __restrict__ slack, double bump, double* __restrict__ w,
                             double* __restrict__ z, double* __restrict__ sc,
                             double* __restrict__ partial) {
    __shared__ double buf[32];
    double mn = INFINITY;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double xi = x[i], lb = lo[i], ub = hi[i];
        const double sl = fmin(cap, fmin(xi - lb, ub - xi));
        slack[i] = sl;
        mn = fmin(mn, sl);
        if (sc) {
            double cd;
            if (isinf(lb) && isinf(ub)) cd = 0.0;
            else if ((xi - lb) < (ub - xi)) cd = fmin(1.0, ub - xi);
            else cd = fmax(-1.0, lb - xi);
            sc[i] = sl * (-cd);
            w[i] += bump;
            z[i] += bump;
        }
    }
    mn = block_reduce(mn, RED_MIN, buf);
    if (threadIdx.x == 0) partial[blockIdx.x] = mn;
}

// g = dg * slack, partials: 0 = max-step(lo,x,hi,g) (:253-266), 1 = sum g^2
__global__ void direction_kernel(size_t n, const double* __restrict__ dg, const double* __restrict__ slack,
                                 const double* __restrict__ x, const double* __restrict__ lo,
                                 const double* __restrict__ hi, double* __restrict__ g,
                                 double* __restrict__ partial) {
    __shared__ double buf[32];
    double mn = INFINITY, ss = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double gi = dg[i] * slack[i];
        g[i] = gi;
        ss = fma(gi, gi, ss);
        if (gi < 0.0) mn = fmin(mn, (lo[i] - x[i]) / gi);
        else if (gi > 0.0) mn = fmin(mn, (hi[i] - x[i]) / gi);
    }
    double v = block_reduce(mn, RED_MIN, buf);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
    v = block_reduce(ss, RED_SUM, buf);
    if (threadIdx.x == 0) partial[gridDim.x + blockIdx.x] = v;
}

// mode 0 (repair, :279-287): step = 0.9*min(maxstep, 1/0.9); x <- max(x + step g, 1e-4)
// mode 1 (recentre, :360-362): step = 0.5*maxstep;           x <- x + step g
// scal[0] = |g|, scal[1] = step
__global__ void step_finish_kernel(const double* __restrict__ partial, int nblocks, int mode,
                                   double* __restrict__ scal) {
    if (threadIdx.x != 0) return;
    double mn = INFINITY, ss = 0.0;
    for (int b = 0; b < nblocks; ++b) {
        mn = fmin(mn, partial[b]);
        ss += partial[nblocks + b];
    }
    const double gamma = 0.9;
    scal[0] = sqrt(ss);
    scal[1] = (mode == 0) ? gamma * fmin(mn, 1.0 / gamma) : 0.5 * mn;
}

__global__ void x_update_kernel(size_t n, int mode, const double* __restrict__ scal,
                                const double* __restrict__ g, double* __restrict__ x) {
    const double step = scal[1];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double v = fma(step, g[i], x[i]);
        x[i] = (mode == 0) ? fmax(v, 1e-4) : v;
    }
}

static int read_scalars(nes_ctx* c, const double* d_src, double* out, int count) {
    NES_TRY(ensure_pinned(c, 64 * sizeof(double)));
    NES_CUDA(c, cudaMemcpyAsync(c->h_pinned, d_src, count * sizeof(double), cudaMemcpyDeviceToHost,
                                c->stream));
    NES_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < count; ++i) out[i] = c->h_pinned[i];
    return 0;
}

}  // namespace nes

extern "C" {

int nes_kkt_newton(nes_matrix* A, nes_factor* L, int filters, const double* l, const double* u,
                   const double* w, const double* z, const double* e, const double* f, const double* g,
                   const double* h, double* dw, double* dx, double* dy, double* dz, nes_ctx* c) {
    NES_ENTER(c);
    if (!A || !l || !u || !w || !z || !e || !f || !g || !h || !dw || !dx || !dy || !dz)
        return fail(c, NES_ERR_INVALID, "nes_kkt_newton: null argument");
    const size_t n = A->base->n, m = A->base->m, pn = pad2(n), pm = pad2(m);
    double* ws = ensure_ws(c, WS_DRIVER, (21 * pn + 2 * pm + 16) * sizeof(double));
    if (!ws) return c->status;
    double* p = ws;
    auto take = [&](size_t len) {
        double* q = p;
        p += len;
        return q;
    };
    double *dl = take(pn), *du = take(pn), *dwv = take(pn), *dzv = take(pn), *de = take(pn),
           *df = take(pn), *dh = take(pn);
    KktVectors kv;
    kv.l = dl; kv.u = du; kv.w = dwv; kv.z = dzv; kv.e = de; kv.f = df; kv.h = dh;
    kv.wp = take(pn); kv.ep = take(pn); kv.lp = take(pn); kv.fp = take(pn); kv.d = take(pn);
    kv.h3 = take(pn); kv.s = take(pn); kv.v = take(pn); kv.r = take(pn);
    kv.dw = take(pn); kv.dx = take(pn); kv.dz = take(pn);
    kv.g = take(pm); kv.dy = take(pm);
    kv.div0 = reinterpret_cast<int*>(take(16));
    const double* hv[7] = {l, u, w, z, e, f, h};
    double* dv[7] = {dl, du, dwv, dzv, de, df, dh};
    for (int k = 0; k < 7; ++k)
        NES_CUDA(c, cudaMemcpyAsync(dv[k], hv[k], n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NES_CUDA(c, cudaMemcpyAsync(kv.g, g, m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NES_CUDA(c, cudaStreamSynchronize(c->stream));

    // solve-delta-y works on a scaled copy of A (cholmod_copy_sparse + scale-sparse!,
    // sparse-newton-solve.lisp:121-126).  The copy is a view: shared values, its own column scale
    // carved out of the workspace (no allocation per call).  Analysis per call unless the caller
    // recycles a factor.
    nes_matrix view;
    view.set_base(A->base);
    view.d_scale = take(pn);
    view.d_theta = take(pn);
    nes_matrix* As = &view;
    nes_factor* Lown = nullptr;
    if (!L) {
        Lown = nes_analyze(As, c);
        if (!Lown) return c->status;
        L = Lown;
    }
    int rc = kkt_newton_dev(c, As, L, filters, kv, nullptr, nullptr);
    if (rc == 0) {
        cudaMemcpyAsync(dw, kv.dw, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(dx, kv.dx, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(dz, kv.dz, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(dy, kv.dy, m * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess)
            rc = fail(c, NES_ERR_CUDA, "nes_kkt_newton: copy back failed");
    }
    if (Lown) nes_free_factor(&Lown, c);
    if (rc > 0) c->status = rc;
    return rc;
}

nes_pdas* nes_pdas_create(nes_matrix* A, const double* cvec, const double* b, const double* lo,
                          const double* hi, const double* x, const double* y, const double* w,
                          const double* z, int filters, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (!A || !cvec || !b || !lo || !hi || !x || !y || !w || !z) {
        fail(c, NES_ERR_INVALID, "nes_pdas_create: null argument");
        return nullptr;
    }
    nes_pdas* st = new nes_pdas();
    st->n = A->base->n;
    st->m = A->base->m;
    st->filters = filters;
    const size_t n = st->n, m = st->m, pn = pad2(n), pm = pad2(m);
    const int gmax = c->num_sms * 2;
    const size_t total = 26 * pn + 5 * pm + (size_t)RED_MAXN * gmax * 2 + 16;
    st->d_block = static_cast<double*>(dev_alloc(c, total * sizeof(double)));
    if (!st->d_block) {
        delete st;
        return nullptr;
    }
    cudaMemsetAsync(st->d_block, 0, total * sizeof(double), c->stream);
    double* p = st->d_block;
    auto take = [&](size_t len) {
        double* q = p;
        p += len;
        return q;
    };
    st->c = take(pn); st->lo = take(pn); st->hi = take(pn); st->x = take(pn); st->w = take(pn);
    st->z = take(pn); st->l = take(pn); st->u = take(pn); st->wu = take(pn); st->zl = take(pn);
    st->rd = take(pn); st->wp = take(pn); st->ep = take(pn); st->lp = take(pn); st->fp = take(pn);
    st->d = take(pn); st->h3 = take(pn); st->s = take(pn); st->v = take(pn); st->r = take(pn);
    st->dw = take(pn); st->dx = take(pn); st->dz = take(pn); st->slack = take(pn); st->gdir = take(pn);
    st->tn = take(pn);
    st->b = take(pm); st->y = take(pm); st->rp = take(pm); st->dy = take(pm); st->tm = take(pm);
    st->partial = take((size_t)RED_MAXN * gmax * 2);
    st->scal = take(16);
    const double* hn[6] = {cvec, lo, hi, x, w, z};
    double* dn[6] = {st->c, st->lo, st->hi, st->x, st->w, st->z};
    bool ok = true;
    for (int k = 0; k < 6 && ok; ++k)
        ok = cudaMemcpyAsync(dn[k], hn[k], n * sizeof(double), cudaMemcpyHostToDevice, c->stream) ==
             cudaSuccess;
    ok = ok && cudaMemcpyAsync(st->b, b, m * sizeof(double), cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(st->y, y, m * sizeof(double), cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
    if (ok) {
        st->A = nes_copy_matrix(A, c);
        if (st->A) st->L = nes_analyze(st->A, c);  // symbolic analysis once; pattern never changes
    }
    if (!ok || !st->A || !st->L) {
        if (ok == false) fail(c, NES_ERR_CUDA, "nes_pdas_create: upload failed");
        nes_pdas_free(&st, c);
        return nullptr;
    }
    return st;
}

int nes_pdas_free(nes_pdas** st, nes_ctx* c) {
    if (!c) return 0;
    if (!st || !*st) return 1;
    if (c->started) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
    }
    nes_free_factor(&(*st)->L, c);
    nes_free_matrix(&(*st)->A, c);
    dev_free(c, (*st)->d_block);
    delete *st;
    *st = nullptr;
    return 1;
}

int nes_pdas_violation(nes_pdas* st, double out[8], nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_pdas_violation: null argument");
    const size_t n = st->n, m = st->m;
    // A x and A' y (sparse-m*, primal-dual-affine-scaling.lisp:145-148)
    NES_TRY(matvec_unscaled(c, st->A->base, 0, 1.0, st->x, 0.0, st->rp));
    NES_TRY(matvec_unscaled(c, st->A->base, 1, 1.0, st->y, 0.0, st->tn));
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        const int gn = vec_grid(c, n), gm = vec_grid(c, m);
        double* pn = st->partial;
        double* pm = st->partial + (size_t)RED_MAXN * c->num_sms * 2;
        pdas_violation_n_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, st->x, st->lo, st->hi, st->w, st->z,
                                                                  st->c, st->tn, st->l, st->u, st->wu,
                                                                  st->zl, st->rd, pn);
        NES_CHECK_LAUNCH(c);
        pdas_violation_m_kernel<<<gm, RED_THREADS, 0, c->stream>>>(m, st->rp, st->b, st->y, pm);
        NES_CHECK_LAUNCH(c);
        pdas_violation_finish_kernel<<<1, 32, 0, c->stream>>>(pn, gn, pm, gm, st->scal);
        NES_CHECK_LAUNCH(c);
    }
    st->have_direction = 0;
    return read_scalars(c, st->scal, out, 8);
}

int nes_pdas_newton_direction(nes_pdas* st, double* step, nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !step) return fail(c, NES_ERR_INVALID, "nes_pdas_newton_direction: null argument");
    KktVectors kv;
    kv.l = st->l; kv.u = st->u; kv.w = st->w; kv.z = st->z; kv.e = st->wu; kv.f = st->zl; kv.h = st->rd;
    kv.g = st->rp;
    kv.wp = st->wp; kv.ep = st->ep; kv.lp = st->lp; kv.fp = st->fp; kv.d = st->d; kv.h3 = st->h3;
    kv.s = st->s; kv.v = st->v; kv.r = st->r;
    kv.dw = st->dw; kv.dx = st->dx; kv.dz = st->dz; kv.dy = st->dy;
    kv.div0 = reinterpret_cast<int*>(st->scal + 15);
    // g2 overwrites st->rp: selector 'p' (Ax - b) is valid between nes_pdas_violation and this call only
    int blocks = 0;
    const int rc = kkt_newton_dev(c, st->A, st->L, st->filters, kv, st->partial, &blocks);
    if (rc != 0) return rc;
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        reduce_finish_kernel<<<1, 32, 0, c->stream>>>(st->partial, blocks, 1, pack_ops(RED_MIN), st->scal);
        NES_CHECK_LAUNCH(c);
    }
    st->have_direction = 1;
    return read_scalars(c, st->scal, step, 1);
}

int nes_pdas_apply_step(nes_pdas* st, double alpha, nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !st->have_direction) return fail(c, NES_ERR_INVALID, "nes_pdas_apply_step: no direction");
    StageTimer t(c, NES_STAGE_VECTOR);
    apply_step_kernel<<<vec_grid(c, st->n > st->m ? st->n : st->m), RED_THREADS, 0, c->stream>>>(
        st->n, st->m, alpha, st->dw, st->dx, st->dz, st->dy, st->w, st->x, st->z, st->y);
    NES_CHECK_LAUNCH(c);
    st->have_direction = 0;  // a direction is applied once (like nes_affine_apply)
    return 0;
}

// shared tail of repair / recentre: given dg in st->tn, g = dg*slack, step rule, x update
static int finish_primal_step(nes_pdas* st, int mode, double out[2], nes_ctx* c) {
    const size_t n = st->n;
    const int gn = vec_grid(c, n);
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        direction_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, st->tn, st->slack, st->x, st->lo, st->hi,
                                                           st->gdir, st->partial);
        NES_CHECK_LAUNCH(c);
        step_finish_kernel<<<1, 32, 0, c->stream>>>(st->partial, gn, mode, st->scal);
        NES_CHECK_LAUNCH(c);
        x_update_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, mode, st->scal, st->gdir, st->x);
        NES_CHECK_LAUNCH(c);
    }
    st->have_direction = 0;
    return read_scalars(c, st->scal, out, 2);
}

int nes_pdas_repair(nes_pdas* st, double out[2], nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_pdas_repair: null argument");
    const size_t n = st->n, m = st->m;
    const int gn = vec_grid(c, n);
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        slack_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, 1e4, st->x, st->lo, st->hi, st->slack, 0.0,
                                                       nullptr, nullptr, nullptr, st->partial);
        NES_CHECK_LAUNCH(c);
    }
    // residual = b - A x  (:248-251)
    NES_CUDA(c, cudaMemcpyAsync(st->tm, st->b, m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    NES_TRY(matvec_unscaled(c, st->A->base, 0, -1.0, st->x, 1.0, st->tm));
    // cholesky-ls! (:223-233): N = A diag(slack); dg = N' (N N')^-1 residual
    NES_TRY(set_scale_dev(c, st->A, st->slack));
    const int rc = factorize_dev(c, st->A, st->L);
    if (rc != 0) return rc;
    NES_TRY(solve_dev(c, st->L, st->tm));
    NES_TRY(matvec(c, st->A, 1, 1.0, st->tm, 0.0, st->tn));
    return finish_primal_step(st, 0, out, c);
}

int nes_pdas_recentre(nes_pdas* st, double out[2], nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_pdas_recentre: null argument");
    const size_t n = st->n;
    const int gn = vec_grid(c, n);
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        // w,z += 1e-4; slack; sc = slack * (-centering-direction)   (:349-357, :305-309)
        slack_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, 1e4, st->x, st->lo, st->hi, st->slack, 1e-4,
                                                       st->w, st->z, st->tn, st->partial);
        NES_CHECK_LAUNCH(c);
    }
    // primal-project (:305-317): A' = A diag(slack); out = sc - A'^T (A' A'^T)^-1 A' sc
    NES_TRY(set_scale_dev(c, st->A, st->slack));
    NES_TRY(matvec(c, st->A, 0, 1.0, st->tn, 0.0, st->tm));
    const int rc = factorize_dev(c, st->A, st->L);
    if (rc != 0) return rc;
    NES_TRY(solve_dev(c, st->L, st->tm));
    NES_TRY(matvec(c, st->A, 1, -1.0, st->tm, 1.0, st->tn));
    return finish_primal_step(st, 1, out, c);
}

int nes_pdas_one_iteration(nes_pdas* st, int repair, double out[9], nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_pdas_one_iteration: null argument");
    double v[8];
    int rc = nes_pdas_violation(st, v, c);
    if (rc != 0) return rc;
    // (assert (every #'plusp l/u)) (:323-324)
    if (!(v[6] > 0.0) || !(v[7] > 0.0))
        return fail(c, NES_ERR_INVALID, "pdas: iterate left the box (min l = %g, min u = %g)", v[6], v[7]);
    const double pobj = v[0], dobj = v[1];
    // fmax/fmin reductions drop NaN, the sums do not: a NaN anywhere in x, y, z, w shows in pobj / dobj
    if (!std::isfinite(pobj) || !std::isfinite(dobj) || !std::isfinite(v[2]) || !std::isfinite(v[3]))
        return fail(c, NES_ERR_INVALID, "pdas: non-finite iterate (pobj = %g, dobj = %g)", pobj, dobj);
    const double gap = fabs(pobj - dobj) / fmax(fmax(fabs(pobj), fabs(dobj)), 1.0);  // :345-346
    out[0] = gap;
    out[1] = dobj;
    out[2] = std::numeric_limits<double>::quiet_NaN();
    out[3] = pobj;
    out[4] = v[2]; out[5] = v[3]; out[6] = v[4]; out[7] = v[5];
    const bool primal_feasible = v[2] < 1e-2;  // :333
    double s2[2];
    if (!primal_feasible) {
        out[8] = 1;
        return nes_pdas_repair(st, s2, c);
    }
    if (repair) {
        out[8] = 2;
        return nes_pdas_recentre(st, s2, c);
    }
    out[8] = 0;
    double step = 0.0;
    rc = nes_pdas_newton_direction(st, &step, c);
    if (rc != 0) return rc;
    const double alpha = fmin(1.0, 0.9 * step);  // :377-378, must lie in (0, 1]
    if (!(alpha > 0.0)) return fail(c, NES_ERR_INVALID, "pdas: step %g is not in (0,1]", alpha);
    rc = nes_pdas_apply_step(st, alpha, c);
    out[2] = step;
    return rc;
}

int nes_pdas_solve(nes_pdas* st, int max_iter, int* iters, double* obj, double* gap, nes_ctx* c) {
    NES_ENTER(c);
    if (!st) return fail(c, NES_ERR_INVALID, "nes_pdas_solve: null state");
    int repair = 0;
    double out[9];
    int i = 1;
    for (; max_iter <= 0 || i <= max_iter; ++i) {
        const int rc = nes_pdas_one_iteration(st, repair, out, c);
        if (rc != 0) {
            if (iters) *iters = i;
            return rc;
        }
        repair = (!std::isnan(out[2]) && out[2] < 1e-6) ? 1 : 0;  // :393
        if (out[0] < 1e-4) {                                        // :394
            if (iters) *iters = i;
            if (obj) *obj = out[1];
            if (gap) *gap = out[0];
            return 0;
        }
    }
    // max_iter exhausted: the Lisp loop falls through and returns NIL (:388-396)
    if (iters) *iters = i - 1;
    if (obj) *obj = out[1];
    if (gap) *gap = out[0];
    c->status = NES_MAXITER;
    return NES_MAXITER;
}

static double* pdas_vec(nes_pdas* st, int which, size_t* len) {
    switch (which) {
        case 'x': *len = st->n; return st->x;
        case 'w': *len = st->n; return st->w;
        case 'z': *len = st->n; return st->z;
        case 'y': *len = st->m; return st->y;
        case 'X': *len = st->n; return st->dx;
        case 'W': *len = st->n; return st->dw;
        case 'Z': *len = st->n; return st->dz;
        case 'Y': *len = st->m; return st->dy;
        case 'l': *len = st->n; return st->l;
        case 'u': *len = st->n; return st->u;
        case 'p': *len = st->m; return st->rp;
        case 'd': *len = st->n; return st->rd;
        case 't': *len = st->n; return st->s;  // sqrt(theta) of the last Newton step
        default: return nullptr;
    }
}

int nes_pdas_get(nes_pdas* st, int which, double* out, nes_ctx* c) {
    NES_ENTER(c);
    size_t len = 0;
    double* p = st ? pdas_vec(st, which, &len) : nullptr;
    if (!p || !out) return fail(c, NES_ERR_INVALID, "nes_pdas_get: bad selector");
    return download(c, out, p, len * sizeof(double));
}

int nes_pdas_set(nes_pdas* st, int which, const double* in, nes_ctx* c) {
    NES_ENTER(c);
    size_t len = 0;
    double* p = st ? pdas_vec(st, which, &len) : nullptr;
    if (!p || !in) return fail(c, NES_ERR_INVALID, "nes_pdas_set: bad selector");
    st->have_direction = 0;  // a direction computed for the previous iterate must not be applied to this one
    return upload(c, p, in, len * sizeof(double));
}

}  // extern "C"

// =================================================================================================
// Primal affine scaling state (affine-scaling.lisp), device resident
//   nes_affine_residual        residual :209-213 (+ c'x)
//   nes_affine_repair          one-repair-iteration :226-243, cholesky-ls! :215-221
//   nes_affine_direction       slack :137-148, centering-direction :150-163, project :98-116,
//                              g = dg*slack, max-step :120-133 and the norms of :182-186
//   nes_affine_apply           x <- x + step g (:205-206)
//   nes_affine_one_iteration   one-iteration :245-263 + one-affine-scaling-iteration :165-207
//   nes_affine_solve           affine-scaling :265-297 (symbolic analysis once, numeric per iteration)
// =================================================================================================
struct nes_affine {
    nes_matrix* A = nullptr;
    nes_factor* L = nullptr;
    size_t n = 0, m = 0;
    double* d_block = nullptr;
    double *c, *l, *u, *x, *slack, *sc, *g;  // n
    double *b, *r, *t;                       // m
    double* partial;
    double* scal;
    int have_direction = 0;
};

namespace nes {

// slack = min(cap, x-l, u-x); sc = slack * (-(centering ? centering-direction : c)); partial: min slack
__global__ void affine_slack_kernel(size_t n, double cap, int centering, const double* __restrict__ x,
                                    const double* __restrict__ lo, const double* __restrict__ hi,
                                    const double* __restrict__ cvec, double* __restrict__ slack,
                                    double* __restrict__ sc, double* __restrict__ partial) {
    __shared__ double buf[32];
    double mn = INFINITY;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double xi = x[i], lb = lo[i], ub = hi[i];
        const double sl = fmin(cap, fmin(xi - lb, ub - xi));
        slack[i] = sl;
        mn = fmin(mn, sl);
        if (sc) {
            double d = cvec[i];
            if (centering) {
                if (isinf(lb) && isinf(ub)) d = 0.0;
                else if ((xi - lb) < (ub - xi)) d = fmin(1.0, ub - xi);
                else d = fmax(-1.0, lb - xi);
            }
            sc[i] = sl * (-1.0 * d);
        }
    }
    mn = block_reduce(mn, RED_MIN, buf);
    if (threadIdx.x == 0) partial[blockIdx.x] = mn;
}

// g = dg*slack; partials: 0 max-step(l,x,u,g)  1 sum g^2  2 sum dg^2  3 sum g*c
__global__ void affine_direction_kernel(size_t n, const double* __restrict__ dg,
                                        const double* __restrict__ slack, const double* __restrict__ x,
                                        const double* __restrict__ lo, const double* __restrict__ hi,
                                        const double* __restrict__ cvec, double* __restrict__ g,
                                        double* __restrict__ partial) {
    __shared__ double buf[32];
    double mn = INFINITY, sg = 0.0, sd = 0.0, sgc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const double di = dg[i], gi = di * slack[i];
        g[i] = gi;
        sg = fma(gi, gi, sg);
        sd = fma(di, di, sd);
        sgc = fma(gi, cvec[i], sgc);
        if (gi < 0.0) mn = fmin(mn, (lo[i] - x[i]) / gi);
        else if (gi > 0.0) mn = fmin(mn, (hi[i] - x[i]) / gi);
    }
    double v = block_reduce(mn, RED_MIN, buf);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
    v = block_reduce(sg, RED_SUM, buf);
    if (threadIdx.x == 0) partial[gridDim.x + blockIdx.x] = v;
    v = block_reduce(sd, RED_SUM, buf);
    if (threadIdx.x == 0) partial[2 * gridDim.x + blockIdx.x] = v;
    v = block_reduce(sgc, RED_SUM, buf);
    if (threadIdx.x == 0) partial[3 * gridDim.x + blockIdx.x] = v;
}

// partials: 0 sum a.a (first na entries), 1 sum b.c (nb entries)
__global__ void two_dots_kernel(size_t na, const double* __restrict__ a, size_t nb,
                                const double* __restrict__ b, const double* __restrict__ cvec,
                                double* __restrict__ partial) {
    __shared__ double buf[32];
    double s0 = 0.0, s1 = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < na; i += stride) s0 = fma(a[i], a[i], s0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nb; i += stride) s1 = fma(b[i], cvec[i], s1);
    double v = block_reduce(s0, RED_SUM, buf);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
    v = block_reduce(s1, RED_SUM, buf);
    if (threadIdx.x == 0) partial[gridDim.x + blockIdx.x] = v;
}

__global__ void axpy_kernel(size_t n, double alpha, const double* __restrict__ v, double* __restrict__ x) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        x[i] = fma(alpha, v[i], x[i]);
}

}  // namespace nes

extern "C" {

nes_affine* nes_affine_create(nes_matrix* A, const double* cvec, const double* b, const double* l,
                              const double* u, const double* x, nes_ctx* c) {
    NES_ENTER_PTR(c);
    if (!A || !cvec || !b || !l || !u || !x) {
        fail(c, NES_ERR_INVALID, "nes_affine_create: null argument");
        return nullptr;
    }
    nes_affine* st = new nes_affine();
    st->n = A->base->n;
    st->m = A->base->m;
    const size_t n = st->n, m = st->m, pn = pad2(n), pm = pad2(m);
    const int gmax = c->num_sms * 2;
    const size_t total = 7 * pn + 3 * pm + (size_t)RED_MAXN * gmax + 16;
    st->d_block = static_cast<double*>(dev_alloc(c, total * sizeof(double)));
    if (!st->d_block) {
        delete st;
        return nullptr;
    }
    cudaMemsetAsync(st->d_block, 0, total * sizeof(double), c->stream);
    double* p = st->d_block;
    auto take = [&](size_t len) {
        double* q = p;
        p += len;
        return q;
    };
    st->c = take(pn); st->l = take(pn); st->u = take(pn); st->x = take(pn); st->slack = take(pn);
    st->sc = take(pn); st->g = take(pn);
    st->b = take(pm); st->r = take(pm); st->t = take(pm);
    st->partial = take((size_t)RED_MAXN * gmax);
    st->scal = take(16);
    const double* hn[4] = {cvec, l, u, x};
    double* dn[4] = {st->c, st->l, st->u, st->x};
    bool ok = true;
    for (int k = 0; k < 4 && ok; ++k)
        ok = cudaMemcpyAsync(dn[k], hn[k], n * sizeof(double), cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(st->b, b, m * sizeof(double), cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
    if (ok) {
        st->A = nes_copy_matrix(A, c);                 // affine-A-copy (:29-35): values shared
        if (st->A) st->L = nes_analyze(st->A, c);      // cholmod_analyze once (:270-271)
    }
    if (!ok || !st->A || !st->L) {
        if (!ok) fail(c, NES_ERR_CUDA, "nes_affine_create: upload failed");
        nes_affine_free(&st, c);
        return nullptr;
    }
    return st;
}

int nes_affine_free(nes_affine** st, nes_ctx* c) {
    if (!c) return 0;
    if (!st || !*st) return 1;
    if (c->started) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
    }
    nes_free_factor(&(*st)->L, c);
    nes_free_matrix(&(*st)->A, c);
    dev_free(c, (*st)->d_block);
    delete *st;
    *st = nullptr;
    return 1;
}

int nes_affine_residual(nes_affine* st, double out[2], nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_affine_residual: null argument");
    NES_CUDA(c, cudaMemcpyAsync(st->r, st->b, st->m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    NES_TRY(matvec_unscaled(c, st->A->base, 0, -1.0, st->x, 1.0, st->r));
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        const int g = vec_grid(c, st->n > st->m ? st->n : st->m);
        two_dots_kernel<<<g, RED_THREADS, 0, c->stream>>>(st->m, st->r, st->n, st->x, st->c, st->partial);
        NES_CHECK_LAUNCH(c);
        reduce_finish_kernel<<<1, 32, 0, c->stream>>>(st->partial, g, 2, pack_ops(RED_SUM, RED_SUM), st->scal);
        NES_CHECK_LAUNCH(c);
    }
    NES_TRY(read_scalars(c, st->scal, out, 2));
    out[0] = sqrt(out[0]);
    return 0;
}

// shared by repair and direction: given the rhs in st->t, factorize (A diag slack)(A diag slack)' and
// solve in place.  Returns 0 or NES_NOT_POSDEF.
static int affine_normal_solve(nes_affine* st, nes_ctx* c) {
    NES_TRY(set_scale_dev(c, st->A, st->slack));
    const int rc = factorize_dev(c, st->A, st->L);
    if (rc != 0) return rc;
    return solve_dev(c, st->L, st->t);
}

int nes_affine_repair(nes_affine* st, double out[2], nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_affine_repair: null argument");
    const size_t n = st->n;
    const int gn = vec_grid(c, n);
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        affine_slack_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, sqrt(1e8), 0, st->x, st->l, st->u, st->c,
                                                              st->slack, nullptr, st->partial);
        NES_CHECK_LAUNCH(c);
    }
    // residual was computed by nes_affine_residual into st->r (the Lisp passes it in, :226)
    NES_CUDA(c, cudaMemcpyAsync(st->t, st->r, st->m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    const int rc = affine_normal_solve(st, c);
    if (rc != 0) return rc;
    NES_TRY(matvec(c, st->A, 1, 1.0, st->t, 0.0, st->sc));  // dg = (A slack)' t
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        affine_direction_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, st->sc, st->slack, st->x, st->l,
                                                                  st->u, st->c, st->g, st->partial);
        NES_CHECK_LAUNCH(c);
        reduce_finish_kernel<<<1, 32, 0, c->stream>>>(st->partial, gn, 4,
                                                     pack_ops(RED_MIN, RED_SUM, RED_SUM, RED_SUM), st->scal);
        NES_CHECK_LAUNCH(c);
    }
    double s4[4];
    NES_TRY(read_scalars(c, st->scal, s4, 4));
    const double gamma = 0.9;
    const double step = gamma * fmin(s4[0], 1.0 / gamma);  // :238
    out[0] = sqrt(s4[1]);
    out[1] = step;
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        axpy_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, step, st->g, st->x);
        NES_CHECK_LAUNCH(c);
    }
    st->have_direction = 0;
    return 0;
}

int nes_affine_direction(nes_affine* st, int centering, double out[5], nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_affine_direction: null argument");
    const size_t n = st->n;
    const int gn = vec_grid(c, n);
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        affine_slack_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, 1e8, centering, st->x, st->l, st->u, st->c,
                                                              st->slack, st->sc, st->partial);
        NES_CHECK_LAUNCH(c);
        reduce_finish_kernel<<<1, 32, 0, c->stream>>>(st->partial, gn, 1, pack_ops(RED_MIN), st->scal + 8);
        NES_CHECK_LAUNCH(c);
    }
    // project (:98-116): AD = A diag(slack); y = (AD AD')^-1 AD sc; dg = sc - AD' y
    NES_TRY(set_scale_dev(c, st->A, st->slack));
    NES_TRY(matvec(c, st->A, 0, 1.0, st->sc, 0.0, st->t));
    const int rc = affine_normal_solve(st, c);
    if (rc != 0) return rc;
    NES_TRY(matvec(c, st->A, 1, -1.0, st->t, 1.0, st->sc));
    {
        StageTimer t(c, NES_STAGE_VECTOR);
        affine_direction_kernel<<<gn, RED_THREADS, 0, c->stream>>>(n, st->sc, st->slack, st->x, st->l,
                                                                  st->u, st->c, st->g, st->partial);
        NES_CHECK_LAUNCH(c);
        reduce_finish_kernel<<<1, 32, 0, c->stream>>>(st->partial, gn, 4,
                                                     pack_ops(RED_MIN, RED_SUM, RED_SUM, RED_SUM), st->scal);
        NES_CHECK_LAUNCH(c);
    }
    double s9[9];
    NES_TRY(read_scalars(c, st->scal, s9, 9));
    out[0] = 0.9 * s9[0];   // step = gamma * max-step (:183)
    out[1] = sqrt(s9[1]);   // |g|
    out[2] = sqrt(s9[2]);   // |dg|
    out[3] = s9[3];         // g.c
    out[4] = s9[8];         // min slack (must be > 0, :144)
    st->have_direction = 1;
    return 0;
}

int nes_affine_apply(nes_affine* st, double step, nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !st->have_direction) return fail(c, NES_ERR_INVALID, "nes_affine_apply: no direction");
    StageTimer t(c, NES_STAGE_VECTOR);
    axpy_kernel<<<vec_grid(c, st->n), RED_THREADS, 0, c->stream>>>(st->n, step, st->g, st->x);
    NES_CHECK_LAUNCH(c);
    st->have_direction = 0;
    return 0;
}

// one-affine-scaling-iteration (:165-207).  *cont receives the Lisp's second value.
static int affine_optimize(nes_affine* st, int centering, int* cont, nes_ctx* c) {
    double d[5];
    const int rc = nes_affine_direction(st, centering, d, c);
    if (rc < 0) return rc;
    if (rc != 0) {  // " singular " (:178-181)
        *cont = 0;
        return 0;
    }
    if (!(d[4] > 0.0)) return fail(c, NES_ERR_INVALID, "affine: iterate left the box (min slack %g)", d[4]);
    const double step = d[0], norm_g = d[1], norm_dg = d[2], descent = d[3];
    if (step > 1e10) return fail(c, NES_ERR_INVALID, "Unbounded problem");  // (:187-188)
    if (!centering) {
        const double nd = (double)st->n;
        if (norm_dg < fmin(1e-6, 1e-8 * nd) || descent > 0.0) {
            *cont = 0;
            return 0;
        }
        if (step * norm_g < 1e-6 || descent > 0.0) return affine_optimize(st, 1, cont, c);
    }
    NES_TRY(nes_affine_apply(st, step, c));
    *cont = 1;
    return 0;
}

int nes_affine_one_iteration(nes_affine* st, int centering, double out[4], nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_affine_one_iteration: null argument");
    double r2[2];
    NES_TRY(nes_affine_residual(st, r2, c));
    out[0] = r2[0];  // |b - Ax| before the step
    out[1] = r2[1];  // c'x before the step
    int cont = 1;
    if (r2[0] > 1e-6 * (double)st->m) {  // (:248)
        double rr[2];
        const int rc = nes_affine_repair(st, rr, c);
        if (rc < 0) return rc;
        if (rc != 0) return fail(c, NES_NOT_POSDEF, "affine repair: Cholesky failed"), NES_NOT_POSDEF;
        out[2] = 1;  // branch: repair
    } else {
        NES_TRY(affine_optimize(st, centering, &cont, c));
        out[2] = centering ? 2 : 0;
    }
    out[3] = cont;
    return 0;
}

int nes_affine_solve(nes_affine* st, int max_iter, int* iters, double* obj, double* resnorm, nes_ctx* c) {
    NES_ENTER(c);
    if (!st) return fail(c, NES_ERR_INVALID, "nes_affine_solve: null state");
    double out[4], r2[2] = {0, 0};
    int i = 0;
    bool stopped = false;
    for (; max_iter <= 0 || i < max_iter; ++i) {
        const int rc = nes_affine_one_iteration(st, ((i + 1) % 16) == 0, out, c);  // (:283)
        if (rc != 0) {
            if (iters) *iters = i + 1;
            return rc;
        }
        NES_TRY(nes_affine_residual(st, r2, c));
        if (!(out[3] != 0.0 || r2[0] > 1e-6 * (double)st->m)) {  // (:284-287)
            ++i;
            stopped = true;
            break;
        }
    }
    if (iters) *iters = i;
    if (obj) *obj = r2[1];
    if (resnorm) *resnorm = r2[0];
    if (!stopped) {  // max_iter is an addition (the Lisp loop has none): say that the stop rule never fired
        c->status = NES_MAXITER;
        return NES_MAXITER;
    }
    return 0;
}

int nes_affine_get(nes_affine* st, int which, double* out, nes_ctx* c) {
    NES_ENTER(c);
    if (!st || !out) return fail(c, NES_ERR_INVALID, "nes_affine_get: null argument");
    switch (which) {
        case 'x': return download(c, out, st->x, st->n * sizeof(double));
        case 'g': return download(c, out, st->g, st->n * sizeof(double));
        case 'r': return download(c, out, st->r, st->m * sizeof(double));
        case 's': return download(c, out, st->slack, st->n * sizeof(double));
        default: return fail(c, NES_ERR_INVALID, "nes_affine_get: bad selector");
    }
}

}  // extern "C"
