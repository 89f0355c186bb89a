/* nes.h -- C ABI of the B200-native normal-equations engine (libnes.so).
 *
 * Drop-in boundary for the hot path of pkhuong/cholesky-is-magic's interior-point LP solvers.
 * The reference crosses from Common Lisp into C through sb-alien (sparse-cholesky.lisp:1-342) to
 * SuiteSparse CHOLMOD plus the accessor shim wrapper.c.  This header declares what a maintainer
 * binds instead: the same call shapes (context with get/set accessors, analyze / factorize / solve /
 * sdmult / scale on opaque handles, handle-taking frees that return non-zero on success), with all
 * arithmetic executed by hand-written sm_100a CUDA kernels.  Plain C types only; every pointer
 * argument is a HOST pointer unless its name ends in _dev.  There is no CPU fallback: if no CUDA
 * device is usable nes_start() fails and every later call returns NES_ERR_NO_DEVICE.
 *
 * Each entry point cites the reference interface it replaces (file:line in the reference tree).
 */
#ifndef NES_H
#define NES_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

/* status codes: same sign convention as cholmod_common.status (0 ok, >0 warning, <0 error) */
#define NES_OK 0
#define NES_NOT_POSDEF 1        /* CHOLMOD_NOT_POSDEF: non-positive pivot, see nes_get_minor() */
#define NES_DIV_BY_ZERO 2       /* solve-kkt-newton divided by a zero u or z: filter-Z (sparse-newton-solve.lisp:40-45)
                                   zeroes z for x - lo > 1e7 and scale-Z then divides by it -- SBCL traps there,
                                   here the step is refused with this status (positive: no result, like 1) */
#define NES_MAXITER 3           /* nes_pdas_solve / nes_affine_solve ran out of iterations before the stop rule
                                   fired (the Lisp loops return NIL); iters/obj/gap hold the last iterate */
#define NES_ERR_NO_DEVICE (-1)  /* no usable sm_100 device / library not started */
#define NES_ERR_OUT_OF_MEMORY (-2)
#define NES_ERR_INVALID (-4)    /* CHOLMOD_INVALID: bad argument */
#define NES_ERR_CUDA (-5)       /* CUDA runtime error; text in nes_last_error() */
#define NES_ERR_COMM (-6)       /* NCCL error */

typedef struct nes_ctx nes_ctx;       /* replaces cholmod_common   (sparse-cholesky.lisp:3) */
typedef struct nes_matrix nes_matrix; /* replaces cholmod_sparse* / cholmod_dense* holding A
                                         (sparse-cholesky.lisp:45-63, 162-171) */
typedef struct nes_factor nes_factor; /* replaces cholmod_factor* + the solve2 workspaces of
                                         solve-sparse-state (sparse-cholesky.lisp:143, 479-484) */

/* ---- context lifetime: wrapper.c:8-16, sparse-cholesky.lisp:389-406 (with-cholmod) ---------- */
nes_ctx* nes_allocate(void);            /* cholmod_allocate  (wrapper.c:8)  */
void nes_release(nes_ctx* c);           /* cholmod_release   (wrapper.c:13) */
int nes_start(nes_ctx* c);              /* cholmod_start     (sparse-cholesky.lisp:40): binds the
                                           current CUDA device (nes_set_device first to pick one) */
int nes_finish(nes_ctx* c);             /* cholmod_finish    (sparse-cholesky.lisp:41) */
int nes_defaults(nes_ctx* c);           /* cholmod_defaults  (sparse-cholesky.lisp:42) */
int nes_free_work(nes_ctx* c);          /* cholmod_free_work (sparse-cholesky.lisp:43) */
int nes_set_device(nes_ctx* c, int device);
const char* nes_last_error(const nes_ctx* c);
int nes_version(int version[3]);        /* cholmod_version   (sparse-cholesky.lisp:258) */

/* ---- accessors: the 19 fields of wrapper.c:31-52, same get/set-returns-old contract --------- */
#define NES_DECLARE_ACCESSOR(FIELD, TYPE)              \
    TYPE nes_get_##FIELD(const nes_ctx* c);            \
    TYPE nes_set_##FIELD(nes_ctx* c, TYPE new_value);
NES_DECLARE_ACCESSOR(print, int)
NES_DECLARE_ACCESSOR(print_function, void*)
NES_DECLARE_ACCESSOR(dbound, double)            /* diagonal clamp; 0 = off like cholmod_defaults */
NES_DECLARE_ACCESSOR(supernodal_switch, double)
NES_DECLARE_ACCESSOR(supernodal, int)
NES_DECLARE_ACCESSOR(selected, int)
NES_DECLARE_ACCESSOR(itype, int)                /* 0: int32 indices (sparse-cholesky.lisp:26) */
NES_DECLARE_ACCESSOR(dtype, int)                /* 0: double */
NES_DECLARE_ACCESSOR(status, int)
NES_DECLARE_ACCESSOR(fl, double)                /* factorization flop count from analyze */
NES_DECLARE_ACCESSOR(lnz, double)               /* nnz(L) */
NES_DECLARE_ACCESSOR(anz, double)               /* nnz(tril(A A')) */
NES_DECLARE_ACCESSOR(modfl, double)
NES_DECLARE_ACCESSOR(malloc_count, size_t)      /* live device+host blocks owned by the context */
NES_DECLARE_ACCESSOR(memory_usage, size_t)      /* peak bytes */
NES_DECLARE_ACCESSOR(memory_inuse, size_t)      /* live bytes (leak test, sparse-newton-solve.lisp:255-258) */
NES_DECLARE_ACCESSOR(rowfacfl, double)
NES_DECLARE_ACCESSOR(aatfl, double)             /* flops to form A A' */
NES_DECLARE_ACCESSOR(blas_ok, int)
#undef NES_DECLARE_ACCESSOR
int nes_get_minor(const nes_ctx* c);            /* column of the failed pivot (cholmod_factor.minor) */

/* ---- constraint matrix A (device resident, immutable values + optional column scale) --------- */
/* A nes_matrix is opaque except for its first three words, which mirror the leading fields of
 * cholmod_sparse (sparse-cholesky.lisp:45-48) so that the Lisp's (slot A 'nrow), (slot A 'ncol) and
 * (slot A 'nzmax) (affine-scaling.lisp:34,218; primal-dual-affine-scaling.lisp:226) keep working: */
struct nes_matrix_header {
    size_t nrow, ncol, nzmax;
};
/* make-dense-from-matlisp + cholmod_dense_to_sparse (sparse-cholesky.lisp:346-368, 411-414):
 * A is m x n column-major with leading dimension ld >= m. */
nes_matrix* nes_dense_to_matrix(const double* A, size_t nrow, size_t ncol, size_t ld, nes_ctx* c);
/* make-sparse-from-triplet-vector (sparse-cholesky.lisp:433-459): cholmod_allocate_triplet +
 * cholmod_triplet_to_sparse (duplicates summed) + cholmod_sort. */
nes_matrix* nes_triplet_to_sparse(const int* row, const int* col, const double* val, size_t nnz,
                                  size_t nrow, size_t ncol, nes_ctx* c);
/* compressed-column input (cholmod_sparse layout p/i/x, sorted or not; packed). */
nes_matrix* nes_csc_to_matrix(const int* colptr, const int* rowidx, const double* val, size_t nrow,
                              size_t ncol, nes_ctx* c);
/* the dense synthetic test matrix of newton-solve.lisp:194-195, A = U(0,1) + eye(m,n), generated on
 * the device from a counter-based hash (same values as lpgen.dense_entry on the host). */
nes_matrix* nes_generate_dense(size_t nrow, size_t ncol, unsigned long long seed, nes_ctx* c);
/* cholmod_copy_sparse (sparse-cholesky.lisp:128): shares the immutable values, own column scale. */
nes_matrix* nes_copy_matrix(nes_matrix* A, nes_ctx* c);
/* cholmod_free_sparse (sparse-cholesky.lisp:75): frees *A, stores NULL, returns non-zero on success. */
int nes_free_matrix(nes_matrix** A, nes_ctx* c);
/* cholmod_scale(s, CHOLMOD_COL=2, A) (sparse-cholesky.lisp:461-473): A <- A diag(s).  `scale` must be
 * 2 (columns).  The product is never materialised: s is folded into formation and sdmult. */
int nes_scale(const double* s, int scale, nes_matrix* A, nes_ctx* c);
/* affine-A-copy (affine-scaling.lisp:29-35): restore the pristine values (drop the column scale). */
int nes_unscale(nes_matrix* A, nes_ctx* c);
size_t nes_matrix_nrow(const nes_matrix* A);
size_t nes_matrix_ncol(const nes_matrix* A);
size_t nes_matrix_nnz(const nes_matrix* A);     /* cholmod_nnz (sparse-cholesky.lisp:84) */
int nes_matrix_is_dense(const nes_matrix* A);
/* row scaling of scale-constraints / rescale-sf (primal-dual-affine-scaling.lisp:50-73,
 * standard-form.lisp:107-134): row i *= 1/max_j|a_ij| when that max >= 1e-6; rowscale_out[i]
 * receives the factor (so the caller can scale b alike).  Modifies the shared values in place. */
int nes_scale_rows_maxabs(nes_matrix* A, double* rowscale_out, nes_ctx* c);

/* ---- y <- alpha*op(A)*x + beta*y : cholmod_sdmult via sparse-m* (sparse-cholesky.lisp:567-614) */
int nes_sdmult(nes_matrix* A, int transpose, const double alpha[2], const double beta[2],
               const double* x, double* y, nes_ctx* c);

/* ---- normal equations: analyze / factorize / solve -------------------------------------------
 * With an unsymmetric A (stype 0) CHOLMOD factorizes A*A' (sparse-cholesky.lisp:408); after
 * nes_scale that is (A diag s)(A diag s)' = A diag(s^2) A'. */
nes_factor* nes_analyze(nes_matrix* A, nes_ctx* c);          /* cholmod_analyze   (:261) */
int nes_factorize(nes_matrix* A, nes_factor* L, nes_ctx* c); /* cholmod_factorize (:265); always
                                   returns 1 like CHOLMOD; failure is reported in status (NOT_POSDEF) */
/* cholmod_solve(sys, L, b) (:270): sys must be 0 (CHOLMOD_A).  x and b are nrow doubles. */
int nes_solve(int sys, nes_factor* L, const double* b, double* x, nes_ctx* c);
/* cholmod_solve2 (:276) as used by solve-sparse-recycle (:524-560): same, workspaces live in L. */
int nes_solve2(int sys, nes_factor* L, const double* b, double* x, nes_ctx* c);
int nes_free_factor(nes_factor** L, nes_ctx* c);             /* cholmod_free_factor (:145) */
/* solve-dense (sparse-cholesky.lisp:409-431) in one call: x with (B B') x = b, B m x n column-major.
 * Returns 0 and fills x, or returns NES_NOT_POSDEF (the Lisp returns NIL). */
int nes_solve_dense(const double* B, size_t nrow, size_t ncol, const double* b, double* x,
                    nes_ctx* c);
/* copy the factor back (testing / parity): L is nrow x nrow column-major lower triangular (dense
 * factor; sparse factors are expanded, with the fill-reducing permutation in perm if non-NULL). */
int nes_factor_to_dense(nes_factor* L, double* Lout, size_t ld, int* perm, nes_ctx* c);
/* copy the formed normal matrix A diag(s^2) A' back (lower triangle valid). */
/* Whole-matrix check of the factorization gate on the device (dense only): forms M = A diag(s^2) A' again,
 * subtracts tril(L) tril(L)' on the FP64 tensor cores and returns out = { ||L L' - M||_F / ||M||_F,
 * ||M||_F, ||L L' - M||_F } (norms of the full symmetric matrices).  A must still carry the scale it was
 * factorized with.  Test / bench instrumentation: the reference has no counterpart. */
int nes_factor_residual(nes_matrix* A, nes_factor* L, double out[3], nes_ctx* c);
int nes_normal_matrix_to_dense(nes_matrix* A, double* Mout, size_t ld, nes_ctx* c);

/* ---- solve-kkt-newton (newton-solve.lisp:139-154 dense, sparse-newton-solve.lisp:150-168 sparse)
 * Inputs l,u,w,z,e,f,h are ncol doubles, g is nrow doubles (none is modified; the Lisp destroys its
 * arguments and callers pass copies).  filters != 0 applies filter-U / filter-Z first
 * (sparse-newton-solve.lisp:30-45).  L may be NULL (analyze per call, like solve-sparse-one-shot) or
 * a factor from nes_analyze (symbolic reuse).  Outputs dw,dx,dz (ncol) and dy (nrow).
 * Returns 0, or NES_NOT_POSDEF when the Cholesky fails (the Lisp signals on the NIL). */
int nes_kkt_newton(nes_matrix* A, nes_factor* L, int filters, const double* l, const double* u,
                   const double* w, const double* z, const double* e, const double* f,
                   const double* g, const double* h, double* dw, double* dx, double* dy, double* dz,
                   nes_ctx* c);

/* ---- device-resident primal-dual affine scaling state (primal-dual-affine-scaling.lisp) -------
 * Holds c, b, lo, hi, x, y, w, z and all iteration temporaries on the GPU; only scalars cross PCIe
 * per iteration.  Function names follow the Lisp. */
typedef struct nes_pdas nes_pdas;
/* make-pdas-state (:8-15, :120-133): vectors are copied to the device.  lo/hi already clamped. */
nes_pdas* nes_pdas_create(nes_matrix* A, const double* cvec, const double* b, const double* lo,
                          const double* hi, const double* x, const double* y, const double* w,
                          const double* z, int filters, nes_ctx* c);
int nes_pdas_free(nes_pdas** st, nes_ctx* c);
/* violation (:135-150) + the scalars of one-pdas-iteration (:325-332): out[0..7] =
 * pobj, dobj, |Ax-b|inf, |dual|inf, |w.u|inf, |z.l|inf, min(l), min(u). */
int nes_pdas_violation(nes_pdas* st, double out[8], nes_ctx* c);
/* direction (:152-164): solve-kkt-newton on the current violation vectors, then pdas-step
 * (:194-198) = min(box-step, pos-step w, pos-step z).  *step receives alpha_max (+inf if none). */
int nes_pdas_newton_direction(nes_pdas* st, double* step, nes_ctx* c);
/* apply-step (:200-207): w,x,y,z -= alpha * (dw,dx,dy,dz). */
int nes_pdas_apply_step(nes_pdas* st, double alpha, nes_ctx* c);
/* one-repair-iteration (:268-288).  out[0] = |g|, out[1] = step. */
int nes_pdas_repair(nes_pdas* st, double out[2], nes_ctx* c);
/* the recentring branch of one-pdas-iteration (:348-366): w,z += 1e-4, projected centering step. */
int nes_pdas_recentre(nes_pdas* st, double out[2], nes_ctx* c);
/* one-pdas-iteration (:319-383) entirely in the library: out[0..2] = gap, dobj, step (NaN when the
 * Lisp returns no step), out[3..8] = pobj and the four violations + branch taken (0 newton,
 * 1 repair, 2 recentre). */
int nes_pdas_one_iteration(nes_pdas* st, int repair, double out[9], nes_ctx* c);
/* pdas (:385-396): loop until gap < 1e-4 or max_iter; returns iterations in *iters, (dobj, gap). */
int nes_pdas_solve(nes_pdas* st, int max_iter, int* iters, double* obj, double* gap, nes_ctx* c);
/* read / write state vectors: which = 'x','y','w','z' or direction 'X','Y','W','Z' (dx,dy,dw,dz) */
int nes_pdas_get(nes_pdas* st, int which, double* out, nes_ctx* c);
int nes_pdas_set(nes_pdas* st, int which, const double* in, nes_ctx* c);

/* ---- device-resident primal affine scaling state (affine-scaling.lisp) ---------------------------
 * x and all temporaries live on the GPU; the symbolic analysis is done once at creation
 * (cholmod_analyze in affine-scaling, :270-271) and every iteration refactorizes numerically
 * (solve-sparse-recycle, sparse-cholesky.lisp:524-560). */
typedef struct nes_affine nes_affine;
/* make-affine-scaling-state (:1-10, :52-90): l,u are the (unclamped, already widened) bounds. */
nes_affine* nes_affine_create(nes_matrix* A, const double* cvec, const double* b, const double* l,
                              const double* u, const double* x, nes_ctx* c);
int nes_affine_free(nes_affine** st, nes_ctx* c);
/* residual (:209-213): out[0] = |b - Ax|_2, out[1] = c'x; the vector stays on the device. */
int nes_affine_residual(nes_affine* st, double out[2], nes_ctx* c);
/* one-repair-iteration (:226-243) on the residual of the last nes_affine_residual.
 * out[0] = |g|, out[1] = step.  Returns NES_NOT_POSDEF when the Cholesky fails. */
int nes_affine_repair(nes_affine* st, double out[2], nes_ctx* c);
/* slack, project (:98-116) of c (or of centering-direction when centering != 0), g = dg*slack:
 * out = { gamma*max-step, |g|, |dg|, g.c, min slack }.  Returns NES_NOT_POSDEF for " singular ". */
int nes_affine_direction(nes_affine* st, int centering, double out[5], nes_ctx* c);
int nes_affine_apply(nes_affine* st, double step, nes_ctx* c);   /* x <- x + step g (:205-206) */
/* one-iteration (:245-263): out = { |b-Ax| before, c'x before, branch (0 optimize, 1 repair,
 * 2 recenter), continue flag }. */
int nes_affine_one_iteration(nes_affine* st, int centering, double out[4], nes_ctx* c);
/* affine-scaling (:265-297): returns iterations, c'x and the final residual norm. */
int nes_affine_solve(nes_affine* st, int max_iter, int* iters, double* obj, double* resnorm, nes_ctx* c);
int nes_affine_get(nes_affine* st, int which, double* out, nes_ctx* c);  /* 'x','g','r','s' */

/* ---- batched small dense LPs (BASELINE config 5: 1024 x (m = 256) via affine scaling) -------------
 * The reference would call affine-scaling once per LP; here the LPs advance together and every kernel
 * (formation, potrf, trsm, trsv, GEMV, reductions) covers the whole batch. */
typedef struct nes_batch nes_batch;
/* A_all: B matrices m x n, column-major each, back to back; vectors: B slices of length n (c, l, u, x)
 * or m (b), back to back.  c_all..x_all may be NULL when only nes_batch_normal_solve is used. */
nes_batch* nes_batch_create(const double* A_all, int B, int m, int n, const double* c_all,
                            const double* b_all, const double* l_all, const double* u_all,
                            const double* x_all, nes_ctx* c);
int nes_batch_free(nes_batch** bt, nes_ctx* c);
/* batched solve-dense (sparse-cholesky.lisp:409-431): (A_b diag s_b)(A_b diag s_b)' x_b = rhs_b for every
 * b; s_all (B x n) may be NULL; status[b] = 0 or NES_NOT_POSDEF. */
int nes_batch_normal_solve(nes_batch* bt, const double* s_all, const double* rhs_all, double* x_all,
                           int* status, nes_ctx* c);
/* batched affine-scaling (affine-scaling.lisp:265-297): iters[b] (negative: Cholesky failed in a repair
 * step), obj[b] = c'x, res[b] = |b - Ax|_2. */
int nes_batch_affine_solve(nes_batch* bt, int max_iter, int* iters, double* obj, double* res, nes_ctx* c);
int nes_batch_get_x(nes_batch* bt, double* x_all, nes_ctx* c);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink/NVSwitch --------------------------------
 * The reference has no distributed code at all; these calls are additions.  M and L are distributed
 * 2D block-cyclically over a P x Q grid of the ranks (default 1 x nranks, NES_DIST_GRID=PxQ); each panel is
 * broadcast in row chunks with ncclBroadcast as soon as its rows are solved, so after nes_factorize every
 * rank holds the complete factor and the solves run replicated.  A and all vectors are replicated.  Rank 0 obtains a 128-byte unique id,
 * the launcher (torch.distributed, MPI, ...) ships it to the other ranks, and every rank calls
 * nes_comm_init before the first nes_analyze. */
int nes_comm_unique_id(unsigned char* id128);
int nes_comm_init(nes_ctx* c, int nranks, int rank, const unsigned char* id128);
int nes_comm_finalize(nes_ctx* c);
int nes_comm_rank(const nes_ctx* c);
int nes_comm_nranks(const nes_ctx* c);
/* host-only: the 128x128 tiles of tril(M) owned by `rank` (m rows, `nranks` ranks); returns the count
 * and fills up to `cap` (tile row, tile column) pairs. */
int nes_dist_plan(int m, int nranks, int rank, int* tile_rows, int* tile_cols, int cap);
/* the same for a P x Q grid (rank = p*Q + q), distribution block nbo (0 = default) and panel chunk of
 * chunk_rows rows (0 = default); *nmsgs = broadcasts per factorization, *nroot = those rooted at `rank` */
int nes_dist_plan_grid(int m, int nbo, int P, int Q, int rank, int chunk_rows, int* tile_rows, int* tile_cols,
                       int cap, int* nmsgs, int* nroot);
/* host-only: the broadcasts of the factorization in order, 8 ints each { panel, root, has_diag, row_start, nblocks,
 * bh, stride, dep }: a message is `nblocks` blocks of `bh` rows, `stride` rows apart, of block column `panel`;
 * head_blocks < 0 = default.  Returns the count, fills up to `cap`. */
int nes_dist_plan_msgs(int m, int nbo, int P, int Q, int chunk_rows, int head_blocks, int* out, int cap);
/* choose the P x Q grid (P*Q = nranks; same call on every rank, before nes_analyze of the factors that use it) */
int nes_dist_set_grid(nes_ctx* c, int P, int Q);
/* distribution block and P x Q process grid the dense factorization of an m x m matrix uses on c's ranks */
int nes_dist_layout(const nes_ctx* c, int m, int* nbo, int* P, int* Q);

/* ---- first-order solver of approx.lisp (APPROX on the penalised primal-dual LP) ---------------------
 * K: sparse matrix with one row per `quadratic` constraint (approx.lisp:36-58, built by make-approx
 * :195-299) over the stacked variables [x | y | z | w]; rhs per row, lin / l / u per variable;
 * complementarity constraints (:85-95) as parallel arrays (x index, y index, x0, flipped).  `scale` != 0
 * applies scale-quadratic (:70-74).  One quadratic may be passed as a dense coefficient vector instead of a
 * row of K (dense_row, dense_rhs; NULL = none): make-approx's duality-gap row touches every variable.
 * The handle borrows K (free it after the handle). */
typedef struct nes_approx nes_approx;
nes_approx* nes_approx_create(nes_matrix* K, const double* rhs, const double* lin, const double* l,
                              const double* u, const int* comp_x, const int* comp_y, const double* comp_x0,
                              const int* comp_flipped, int ncomp, const double* dense_row, double dense_rhs,
                              int scale, double z0, nes_ctx* c);
int nes_approx_free(nes_approx** st, nes_ctx* c);
/* value-&-gradient (approx.lisp:338-351) at a host vector x: sum of constraint values, gradient (may be
 * NULL), max |constraint value| */
int nes_approx_value_gradient(nes_approx* st, const double* x, double* value, double* g, double* maxv,
                              nes_ctx* c);
/* which: 'n' nu (accumulate-nu, :97-113), 's' row scales, 'd' {scale, beta} of the dense row,
 * 'z' / 'x' current iterates */
int nes_approx_get(nes_approx* st, int which, double* out, nes_ctx* c);
/* approx (approx.lisp:425-459): up to n_iter iterations from x0 (NULL = 0, projected on the bounds);
 * stops when the projected-gradient norm drops below 1e-10.  stats[7] = {|g|, projected gradient,
 * max constraint value, value + z0, last g.(zp - z), theta, z0 + value of the linear term}. */
int nes_approx_solve(nes_approx* st, int n_iter, const double* x0, double* z_out, int* iters, int* restarts,
                     double* stats, nes_ctx* c);

/* variant 1 = the inner solver of alm-approx.lisp (:198-346): step damped by 0.95, stop when i > 10 and the
 * projected gradient is below `accuracy` (or at the last iteration), max ignores the linear term. */
int nes_approx_set_variant(nes_approx* st, int variant, double accuracy, nes_ctx* c);
/* make-alm-subproblem (alm-approx.lisp:355-403) over the same rows: uniform row scale sqrt(weight), linear
 * term c + A'lambda, constant z0 = -lambda.b; nu is recomputed on the device. */
int nes_approx_set_subproblem(nes_approx* st, double row_scale, const double* lin, double z0, nes_ctx* c);

/* ---- symbolic analysis on its own (host only, no device needed) ---------------------------------
 * What nes_analyze computes for a sparse A before anything touches the GPU: the fill-reducing
 * ordering (nested dissection + reverse Cuthill-McKee leaves), supernodes, assembly-tree levels, index
 * maps and the subtree-to-rank mapping of a multi-GPU run.  Exposed so hosts and tests can inspect or
 * cache the analysis; named int arrays ("perm", "first", "nr", "ld", "rows", "rowptr", "sparent", "level",
 * "lvlptr", "childptr", "child", "relptr", "rel", "cut", "ldu", "owner", "ei", "ej"), long long arrays
 * ("off", "uoff", "edest") and double scalars ("anz", "aatfl", "lnz", "fl", "lsize", "usize", "nsuper",
 * "nlevels").  cholmod_analyze (sparse-cholesky.lisp:261, 509). */
void* nes_symbolic_create(int nrow, int ncol, const int* colptr, const int* rowidx, int nranks,
                          int nd_leaf, char* err, size_t errlen);
/* returns the element count (or -1 for an unknown name) and stores the array's address in *data */
long long nes_symbolic_ints(const void* sym, const char* name, const int** data);
long long nes_symbolic_longs(const void* sym, const char* name, const long long** data);
double nes_symbolic_scalar(const void* sym, const char* name);
void nes_symbolic_free(void* sym);
/* nested dissection stops at vertex sets of `leaf` rows (0 = default max(256, nrow/128)); applies to
 * later nes_analyze calls on this context; returns the previous value */
int nes_set_ordering_leaf(nes_ctx* c, int leaf);

/* ---- instrumentation (bench.py / roofline): device time of the library's own stages ---------- */
#define NES_STAGE_FORM 0      /* K1 fused scale+SYRK   */
#define NES_STAGE_FACTOR 1    /* K2 Cholesky            */
#define NES_STAGE_SOLVE 2     /* K5 triangular solves   */
#define NES_STAGE_GEMV 3      /* K6 GEMV / SpMV         */
#define NES_STAGE_VECTOR 4    /* K7/K8 elementwise + reductions */
#define NES_NUM_STAGES 5
int nes_timing_enable(nes_ctx* c, int on);            /* CUDA events on the library's stream */
int nes_timing_reset(nes_ctx* c);
/* accumulated milliseconds and number of timed intervals for a stage since the last reset */
int nes_timing_get(nes_ctx* c, int stage, double* ms, long long* count);
long long nes_get_launch_count(const nes_ctx* c);     /* kernels launched by this context */
/* algorithmic flops of the last dense formation launch timed as NES_STAGE_FORM: m^2 n, minus the d^2 n of
 * the last d columns of M when their formation is deferred into the factorization stage (dense_chol.cu) */
double nes_get_form_flops(const nes_ctx* c);
/* whole-region device timing: nes_mark_begin records a CUDA event on the library's stream,
 * nes_mark_end records a second one, waits for it and returns the elapsed milliseconds. */
int nes_mark_begin(nes_ctx* c);
int nes_mark_end(nes_ctx* c, double* ms);
int nes_synchronize(nes_ctx* c);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* NES_H */
