// Internal structures of libnes.so.  Public contract: include/nes.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/nes.h"

// ---- context (plays cholmod_common; wrapper.c:18-52 lists the fields the Lisp touches) ----------
struct nes_ctx {
    // wrapper.c accessor fields
    int print = 3;
    void* print_function = nullptr;
    double dbound = 0.0;
    double supernodal_switch = 40.0;
    int supernodal = 1;  // CHOLMOD_AUTO
    int selected = 0;
    int itype = 0;
    int dtype = 0;
    int status = 0;
    double fl = 0, lnz = 0, anz = 0, modfl = 0;
    size_t malloc_count = 0, memory_usage = 0, memory_inuse = 0;
    double rowfacfl = 0, aatfl = 0;
    int blas_ok = 1;
    int minor = -1;
    int nd_leaf = 0;  // nested-dissection leaf size for sparse analysis (0 = default)

    // runtime
    int started = 0;
    int device = -1;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    // low-priority side stream + events for the look-ahead of the dense Cholesky (trailing update of
    // panel J overlaps the factorization of panel J+1, which runs on `stream` at high priority)
    cudaStream_t stream_aux = nullptr;
    cudaEvent_t ev_panel = nullptr, ev_update = nullptr, ev_aux = nullptr;
    // second high-priority stream: the narrow supernodes of a sparse level (potrf -> trsm) run beside the
    // wide ones instead of in front of them (sparse_chol.cu: run_factor_phase)
    cudaStream_t stream_b = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // medium-priority stream: the bulk row pieces of a distributed panel (dense_chol.cu)
    cudaStream_t stream_c = nullptr;
    char err[512] = {0};
    long long launches = 0;
    double form_flops = 0;  // algorithmic flops of the last up-front formation launch (nes_get_form_flops)

    // multi-GPU (one process per GPU; P x Q block-cyclic distribution of M, see nes_dist.cu); rank = p*Q + q
    int nranks = 1, rank = 0;
    int grid_p = 1, grid_q = 1;
    void* nccl_comm = nullptr;

    // allocation ledger (device + pinned host), so frees can be accounted without sizes
    std::unordered_map<void*, size_t> ledger;

    // workspaces owned by the context ("Common workspace", released by nes_free_work / nes_finish).
    // One slot per internal user so nested stages never alias each other's scratch.
    static constexpr int kNumWs = 5;
    double* d_ws[kNumWs] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t ws_bytes[kNumWs] = {0, 0, 0, 0, 0};
    int* d_split_counters = nullptr;  // tail-split arrival counters of dmma_nt (zero between launches)
    double* h_pinned = nullptr;  // small pinned staging buffer for scalar read-backs
    size_t pinned_bytes = 0;

    // stage timing
    int timing = 0;
    struct Interval {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Interval> intervals;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t mark_a = nullptr, mark_b = nullptr;
    double stage_ms[NES_NUM_STAGES] = {0};
    long long stage_count[NES_NUM_STAGES] = {0};
};

namespace nes {

int fail(nes_ctx* c, int status, const char* fmt, ...);
#define NES_CUDA(c, call)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return nes::fail((c), NES_ERR_CUDA, "%s failed: %s (%s:%d)", #call,            \
                             cudaGetErrorString(e__), __FILE__, __LINE__);                 \
    } while (0)
#define NES_CHECK_LAUNCH(c)                                                                \
    do {                                                                                   \
        ++(c)->launches;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess)                                                            \
            return nes::fail((c), NES_ERR_CUDA, "kernel launch failed: %s (%s:%d)",        \
                             cudaGetErrorString(e__), __FILE__, __LINE__);                 \
    } while (0)
#define NES_TRY(expr)                \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ < 0) return rc__;   \
    } while (0)

#define NES_ENTER_PTR(c)                                                 \
    do {                                                                 \
        if (!(c) || !(c)->started) {                                     \
            if (c) nes::fail((c), NES_ERR_NO_DEVICE, "context not started"); \
            return nullptr;                                              \
        }                                                                \
        cudaSetDevice((c)->device);                                      \
    } while (0)
#define NES_ENTER(c)                                                                        \
    do {                                                                                    \
        if (!(c)) return NES_ERR_INVALID;                                                   \
        if (!(c)->started) return nes::fail((c), NES_ERR_NO_DEVICE, "context not started"); \
        cudaSetDevice((c)->device);                                                         \
    } while (0)

void* dev_alloc(nes_ctx* c, size_t bytes);  // nullptr on failure (status set)
void dev_free(nes_ctx* c, void* p);
void* pinned_alloc(nes_ctx* c, size_t bytes);
void pinned_free(nes_ctx* c, void* p);
enum { WS_API = 0, WS_MATVEC = 1, WS_REDUCE = 2, WS_DRIVER = 3, WS_SPLIT = 4 };
// grow-only device workspace slot; returns nullptr on failure (status set)
double* ensure_ws(nes_ctx* c, int slot, size_t bytes);
int ensure_pinned(nes_ctx* c, size_t bytes);
int upload(nes_ctx* c, void* dst_dev, const void* src_host, size_t bytes);
int download(nes_ctx* c, void* dst_host, const void* src_dev, size_t bytes);

// RAII stage timer: records CUDA events on the context's stream around a library stage.
struct StageTimer {
    nes_ctx* c;
    int idx;
    StageTimer(nes_ctx* c, int stage);
    ~StageTimer();
};

// ---- matrices ----------------------------------------------------------------------------------
struct MatrixBase {
    int refs = 1;
    bool dense = true;
    size_t m = 0, n = 0;
    // dense: column-major, ld multiple of 16 doubles (columns 128B aligned)
    size_t ld = 0;
    double* d_val = nullptr;
    CUtensorMap map;  // operand map for dmma_nt (dense only)
    // sparse: CSC (sorted, packed, duplicates summed), int32 indices like the reference's itype 0
    size_t nnz = 0;
    int* d_colptr = nullptr;
    int* d_rowidx = nullptr;
    double* d_values = nullptr;
    // CSR mirror (row-gather SpMV without atomics => deterministic): values are re-gathered
    // from the CSC array through d_csr_src whenever the CSC values change (row scaling)
    int* d_rowptr = nullptr;
    int* d_colidx = nullptr;
    int* d_csr_src = nullptr;
    double* d_csr_val = nullptr;
    std::vector<int> h_colptr, h_rowidx;  // host copy of the pattern for symbolic analysis
};

}  // namespace nes

struct nes_matrix {
    // Public prefix (include/nes.h): the leading fields of cholmod_sparse (sparse-cholesky.lisp:45-48), so
    // the Lisp's (slot A 'nrow) / (slot A 'ncol) / (slot A 'nzmax) keep working on a (* nes-matrix).
    size_t nrow = 0, ncol = 0, nzmax = 0;
    nes::MatrixBase* base = nullptr;
    void set_base(nes::MatrixBase* b) {
        base = b;
        nrow = b->m;
        ncol = b->n;
        nzmax = b->dense ? b->m * b->n : b->nnz;
    }
    double* d_scale = nullptr;  // column scale s (n doubles) or nullptr (cholmod_scale folded lazily)
    double* d_theta = nullptr;  // s^2 padded to a multiple of 16 (formation operand), or nullptr
};

namespace nes {
struct SparseFactor;
}

namespace nes {
struct DistPlan;
}

struct nes_factor {
    bool dense = true;
    nes::SparseFactor* sparse = nullptr;  // supernodal factor (sparse_chol.cu) when !dense
    size_t m = 0;
    size_t ld = 0;
    double* d_M = nullptr;     // m x m column-major: lower triangle holds M, then L in place
    double* d_dinv = nullptr;  // 1/L_jj
    double* d_rhs = nullptr;   // solve workspace (solve2's Y/E)
    int* d_info = nullptr;     // {status, minor}
    int* d_flags = nullptr;    // per-block-row ready flags of the dataflow TRSV (epoch numbered)
    double* d_Winv = nullptr;  // inverses of the 128x128 diagonal blocks of L (solve phase only)
    int flag_epoch = 0;
    CUtensorMap mapM;    // 132 x 32 operand boxes (dmma_nt)
    CUtensorMap mapM68;  // 68 x 32 boxes: the column operand of the half-tile update kernel (dmma_nt64)
    CUtensorMap mapBlk;  // 128 x 128 block boxes (diagonal-block kernels)
    CUtensorMap mapSlab; // 64 x 128 slabs of a panel (TRSM)
    int factorized = 0;
    nes_matrix* analyzed_for = nullptr;
    // pre-scaled operand of the formation: As = A diag(s), refreshed by one elementwise pass per factorization
    // (dense_chol.cu: dense_form_normal); owned by the factor, which outlives the per-call matrix views
    double* d_As = nullptr;
    size_t As_ld = 0, As_n = 0;
    CUtensorMap mapAs;
    // distributed factorization (nranks > 1): tiles of the lower triangle owned by this rank, ordered by
    // outer block column; tile_first[J] = index of the first owned tile whose block column is > J
    int nbo = 0;                 // distribution block = outer panel width
    int2* d_tile_list = nullptr;
    int ntiles_owned = 0;
    nes::DistPlan* dist = nullptr;  // message schedule, tile segments, events, staging ring (nes_dist.cu)
    // deferred formation (single GPU, dense_chol.cu): the tiles of M in block columns >= defer_split are
    // not formed up front but strip by strip inside the factorization, where they fill the SMs the
    // panel chain leaves idle.  Planned once per (m, n); 0 = everything is formed up front.
    struct DeferStrip {
        int col0;    // first column of the strip (a panel boundary)
        int first;   // its tiles: d_defer_tiles[first .. first + ntiles)
        int ntiles;
        int split;   // k-ranges per tile (1 = whole tiles)
    };
    int defer_planned = 0;
    size_t defer_n = 0;
    int defer_split = 0;
    int defer_ntiles_a = 0;          // up-front tiles: d_defer_tiles[0 .. defer_ntiles_a)
    int2* d_defer_tiles = nullptr;
    double* d_defer_ws = nullptr;    // partial tiles of the k-split strips
    int* d_defer_counters = nullptr; // arrival counters (zero between launches)
    std::vector<DeferStrip> defer_strips;
};

namespace nes {

// ---- kernels / stage drivers (one per .cu) ------------------------------------------------------
// K1: M(lower) = A diag(theta) A'  (dense_form.cu)
// allow_defer: leave the block columns >= L->defer_split to dense_cholesky (which then needs A)
int dense_form_normal(nes_ctx* c, const nes_matrix* A, nes_factor* L, bool allow_defer = false);
// K2: in-place blocked Cholesky of L->d_M (dense_chol.cu); sets c->status / c->minor
// A != nullptr: the deferred block columns of M are still to be formed from A (see dense_form_normal)
int dense_cholesky(nes_ctx* c, nes_factor* L, const nes_matrix* A = nullptr);
// K5: x <- (L L')^{-1} x, x device vector of length m (dense_solve.cu)
int dense_solve_inplace(nes_ctx* c, nes_factor* L, double* d_x);
// K6: y <- alpha op(A diag(s)) x + beta y on device vectors (gemv.cu); handles dense and CSC
int matvec(nes_ctx* c, const nes_matrix* A, int transpose, double alpha, const double* d_x,
           double beta, double* d_y);
// batched building blocks (dense_chol.cu / dense_solve.cu), used by batch.cu
int chol_panel_launch(nes_ctx* c, const CUtensorMap& mapBlk, const CUtensorMap& mapSlab, int i0, int ib,
                      int m, double* dinv, int* info, int nbatch, int brows);
int dense_trsv_sweeps(nes_ctx* c, const double* M, long long ld, int m, const double* dinv, double* d_x,
                      int nbatch, int brows);
// multi-GPU pieces (nes_dist.cu)
// One broadcast of a distributed panel: a set of block rows of the panel (all from one root).
struct DistMsg {
    int root = 0;         // world rank that computes and sends it
    int has_diag = 0;     // contains the diagonal block (+ the panel's dinv rides at the end of the message)
    int urgent = 0;       // on the panel chain: its root works on the main (high-priority) stream
    int row_start = 0;    // first row of the first block
    int nblocks = 1;      // blocks of `bh` rows, `stride` rows apart (P x Q: the root's block rows of one chunk)
    int bh = 0, stride = 0;
    int rows = 0;         // nblocks * bh: leading dimension of the packed message (rows past m stay unused)
    int dep = -1;         // index of the last message of the PREVIOUS panel the root's update of these rows needs
    std::vector<int> seg; // root only: tpb + 1 offsets into the rank's tile list, one range per tile column
};
struct DistPanel {
    int j0 = 0, jbo = 0, group_q = 0;
    int diag_msg = 0;            // index of the message that carries the diagonal block
    std::vector<DistMsg> msgs;   // in broadcast order (the same on every rank)
    int col_begin = 0, col_end = 0;  // this rank's tiles of block column J: [col_begin, col_end)
};
struct DistPlan {
    int m = 0, nbo = 0, P = 1, Q = 1, rank = 0, chunk = 0, nblk = 0, tpb = 1;
    std::vector<DistPanel> panels;
    std::vector<int2> tiles;     // this rank's tiles of tril(M), ordered by (block column, message, tile column, row)
    size_t max_msg_doubles = 0;
    // device / stream objects (created by dense_chol.cu on first use, destroyed with the factor)
    static constexpr int kStages = 4;
    double* d_stage[kStages] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> ev_arrived;  // per message, global index (panel-major)
    std::vector<cudaEvent_t> ev_packed;   // per message (used on the root)
    std::vector<cudaEvent_t> ev_colready; // per panel: block column J carries every update up to panel J-2
    std::vector<cudaEvent_t> ev_diagdone; // per panel: the diagonal block is factored (root of the diag message)
    std::vector<int> msg_base;            // global index of panel J's first message
    cudaEvent_t ev_start = nullptr, ev_end[3] = {nullptr, nullptr, nullptr};
};
int dist_make_plan(DistPlan& plan, int m, int nbo, int P, int Q, int rank, int chunk_rows, int head_blocks = -1);
int dist_head_blocks(int nranks);
void dist_free_plan(nes_ctx* c, DistPlan* plan);
int dist_chunk_rows(int m, int nbo, int P);
int dist_broadcast(nes_ctx* c, double* d_buf, size_t count, int root, cudaStream_t stream = nullptr);
int dist_allreduce_int(nes_ctx* c, int* d_buf, size_t count, int op_max_else_min);
int dense_outer_block(int m, int nranks);
int dist_owner(int J, int nranks);
// install a column scale that already lives on the device (nes_scale without the PCIe hop)
int set_scale_dev(nes_ctx* c, nes_matrix* A, const double* d_s);
// plain op(A) product ignoring the column scale
int matvec_unscaled(nes_ctx* c, const MatrixBase* A, int transpose, double alpha, const double* d_x,
                    double beta, double* d_y);

}  // namespace nes
