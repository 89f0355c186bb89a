// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Distribution: the normal matrix M and its factor are laid out 2D block-cyclically in
// nbo x nbo blocks over a P x Q process grid (rank = p*Q + q): block (I, J) of the lower triangle belongs to
// process row I mod P and to the process column that owns block column J (columns are dealt back and forth,
// 0..Q-1, Q-1..0, which balances the triangle).  Default grid: 1 x nranks -- with NVSwitch every GPU receives
// a broadcast at full link bandwidth, so splitting the panel rows as well (P > 1) only adds a latency-bound
// exchange of the diagonal block per panel; NES_DIST_GRID=PxQ selects another grid (measured in DESIGN.md).
//
// A panel (block column J) travels as a fixed sequence of MESSAGES, the same on every rank: [P > 1: the
// diagonal block] and then, for every chunk of `chunk` rows (absolute row ranges, so chunk k of panel J+1
// needs exactly chunk k of panel J) and every process row, the block rows that rank owns.  Each message is
// one ncclBroadcast on a dedicated communication stream, issued as soon as its root has solved those rows:
// the next panel's owner starts on the first chunk while the later ones are still being solved and sent
// (dense_chol.cu has the schedule).  A and the IPM vectors are replicated (A is regenerated or uploaded
// once per rank; all ranks run the O(n) vector kernels redundantly and deterministically, so no scalar
// ever needs to be exchanged).
//
// NCCL is resolved at run time (dlopen) so the library still loads on hosts without it; the unique id
// travels through the caller's launcher (torch.distributed in bench.py / tests).
#include <algorithm>
#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

#include "nes_internal.h"

namespace nes {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    // prefer a copy that is already in the process (torch bundles its own libnccl.so.2)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommInitRankConfig = reinterpret_cast<decltype(api.CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(h, "ncclBroadcast"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.AllReduce)
        return nullptr;
    api.handle = h;
    return &api;
}

// Distribution block = outer panel width of the distributed factorization.  Wide panels (512) make the
// trailing updates efficient, which is what bounds the step on 1-4 GPUs; at 8 GPUs every rank's share of
// the update is small and the panel chain (owner's column update + panel + broadcast) bounds it, and
// that chain is shorter per column with 256-wide panels.  NES_DIST_NBO overrides (testing).
int dense_outer_block(int m, int nranks) {
    if (const char* e = getenv("NES_DIST_NBO")) {
        const int v = atoi(e);
        if (v == 128 || v == 256 || v == 512 || v == 1024) return v;
    }
    if (m <= 12288) return 256;
    return nranks >= 8 ? 256 : 512;  // measured at m = 32768: N=2 206 vs 220 ms, N=4 127 vs 130 ms, N=8 98 vs 94 ms
}

// Owner of outer block column J.  Plain cyclic ownership always hands rank 0 the longest column of
// every round (11% more tiles than average at m = 32768 on 8 ranks); walking the ranks back and forth
// (0..P-1, P-1..0, ...) pairs a long column with a short one and balances the triangle.
int dist_owner(int J, int nranks) {
    const int q = J % (2 * nranks);
    return q < nranks ? q : 2 * nranks - 1 - q;
}

// Rows per chunk of a panel: a multiple of P*nbo (every process row owns whole blocks of every chunk).
// Small chunks shorten the pipeline stages (the next owner starts sooner), large ones cost fewer launches
// and broadcasts.  NES_DIST_CHUNK overrides (testing / tuning).
int dist_chunk_rows(int m, int nbo, int P) {
    int want = 8192;
    if (const char* e = getenv("NES_DIST_CHUNK")) {
        const int v = atoi(e);
        if (v > 0) want = v;
    }
    const int unit = P * nbo;
    int r = (want + unit - 1) / unit * unit;
    if (r < unit) r = unit;
    (void)m;
    return r;
}

// Leading single-block messages of a panel on a 1 x Q grid: block row J (the diagonal block) and block row J+1
// travel on their own, ahead of the row chunks.  The diagonal block of panel J+1 needs nothing else from panel
// J, so the latency chain  potrf(J) -> trsm of block row J+1 -> update + potrf(J+1)  no longer waits for a whole
// chunk to be updated, solved, packed and sent.  NES_DIST_HEAD overrides (0 = the chunks only).
int dist_head_blocks(int nranks) {
    if (const char* e = getenv("NES_DIST_HEAD")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 4) return v;
    }
    (void)nranks;
    return 0;
}

// Message schedule + this rank's tile list.  Host only (checked without a GPU through nes_dist_plan_grid).
int dist_make_plan(DistPlan& pl, int m, int nbo, int P, int Q, int rank, int chunk_rows, int head_blocks) {
    if (head_blocks < 0) head_blocks = dist_head_blocks(P * Q);
    pl = DistPlan();
    pl.m = m; pl.nbo = nbo; pl.P = P; pl.Q = Q; pl.rank = rank; pl.chunk = chunk_rows;
    pl.tpb = nbo / 128;
    pl.nblk = (m + nbo - 1) / nbo;
    const int myp = rank / Q, myq = rank % Q;
    const int tm = (m + 127) / 128, tpb = pl.tpb, nblk = pl.nblk;
    pl.panels.resize(nblk);
    pl.msg_base.assign(nblk + 1, 0);
    for (int J = 0; J < nblk; ++J) {
        DistPanel& pn = pl.panels[J];
        pn.j0 = J * nbo;
        pn.jbo = (m - pn.j0 < nbo) ? m - pn.j0 : nbo;
        pn.group_q = dist_owner(J, Q);
        const int pJ = J % P;
        if (P > 1) {  // the diagonal block goes first, to everyone (the P ranks of the column need L_JJ)
            DistMsg d;
            d.root = pJ * Q + pn.group_q;
            d.has_diag = 1;
            d.urgent = 1;
            d.row_start = pn.j0;
            d.nblocks = 1;
            d.bh = nbo;
            d.stride = 0;
            d.rows = nbo;
            pn.msgs.push_back(d);
        }
        int body0 = pn.j0;  // first row of the chunked part
        if (P == 1)
            for (int hb = 0; hb < head_blocks && body0 < m; ++hb) {
                DistMsg d;
                d.root = pn.group_q;
                d.has_diag = (hb == 0);
                d.urgent = 1;
                d.row_start = body0;
                d.nblocks = 1;
                d.bh = nbo;
                d.stride = 0;
                d.rows = nbo;
                pn.msgs.push_back(d);
                body0 += nbo;
            }
        const int c_first = body0 / chunk_rows, c_last = (m - 1) / chunk_rows;
        for (int ck = c_first; ck <= c_last && body0 < m; ++ck) {
            const int r_lo = std::max(body0, ck * chunk_rows), r_hi = std::min(m, (ck + 1) * chunk_rows);
            if (P == 1) {
                DistMsg d;
                d.root = pn.group_q;
                d.has_diag = (r_lo == pn.j0);
                d.urgent = d.has_diag;
                d.row_start = r_lo;
                d.nblocks = 1;
                d.bh = (r_hi - r_lo + 63) / 64 * 64;
                d.stride = 0;
                d.rows = d.bh;
                pn.msgs.push_back(d);
                continue;
            }
            // block rows I > J inside [r_lo, r_hi), process row by process row, starting below the diagonal
            const int I_lo = std::max(J + 1, r_lo / nbo), I_hi = (r_hi + nbo - 1) / nbo;  // [I_lo, I_hi)
            for (int k = 0; k < P; ++k) {
                const int p = (pJ + 1 + k) % P;
                int I0 = I_lo + ((p - I_lo % P) % P + P) % P;
                if (I0 >= I_hi) continue;
                DistMsg d;
                d.root = p * Q + pn.group_q;
                d.row_start = I0 * nbo;
                d.nblocks = (I_hi - 1 - I0) / P + 1;
                d.bh = nbo;
                d.stride = P * nbo;
                d.rows = d.nblocks * nbo;
                pn.msgs.push_back(d);
            }
        }
        for (size_t k = 0; k < pn.msgs.size(); ++k)
            if (pn.msgs[k].has_diag) pn.diag_msg = (int)k;
        pl.msg_base[J + 1] = pl.msg_base[J] + (int)pn.msgs.size();
        for (const DistMsg& d : pn.msgs)
            pl.max_msg_doubles = std::max(pl.max_msg_doubles, (size_t)d.rows * pn.jbo + (size_t)nbo);
    }
    // which message of the previous panel must have arrived before the root updates a message's rows:
    // the ones that carry the same rows (A operand) and block row J (B operand); the communication stream
    // is in order, so the last of them is enough
    auto rows_of = [&](const DistMsg& d, int b) {  // [lo, hi) of block b of the message
        const int lo = d.row_start + b * d.stride;
        return std::make_pair(lo, std::min(m, lo + d.bh));
    };
    auto overlaps = [&](const DistMsg& d, int lo, int hi) {
        for (int b = 0; b < d.nblocks; ++b) {
            auto r = rows_of(d, b);
            if (r.first < hi && lo < r.second) return true;
        }
        return false;
    };
    for (int J = 1; J < nblk; ++J) {
        DistPanel& pn = pl.panels[J];
        const DistPanel& pv = pl.panels[J - 1];
        for (DistMsg& d : pn.msgs) {
            int dep = -1;
            for (size_t k = 0; k < pv.msgs.size(); ++k) {
                bool need = overlaps(pv.msgs[k], pn.j0, pn.j0 + pn.jbo);
                for (int b = 0; b < d.nblocks && !need; ++b) {
                    auto r = rows_of(d, b);
                    need = overlaps(pv.msgs[k], r.first, r.second);
                }
                if (need) dep = (int)k;
            }
            d.dep = dep;
        }
    }
    // this rank's tiles: by block column, then message, then tile column, then tile row
    for (int J = 0; J < nblk; ++J) {
        DistPanel& pn = pl.panels[J];
        pn.col_begin = (int)pl.tiles.size();
        if (pn.group_q == myq) {
            const int c_lo = J * tpb, c_hi = std::min(tm, c_lo + tpb);
            for (DistMsg& d : pn.msgs) {
                if (d.root != rank) continue;
                d.seg.assign(tpb + 1, (int)pl.tiles.size());
                for (int t = 0; t < tpb; ++t) {
                    d.seg[t] = (int)pl.tiles.size();
                    const int bj = c_lo + t;
                    if (bj < c_hi)
                        for (int b = 0; b < d.nblocks; ++b) {
                            auto r = rows_of(d, b);
                            for (int bi = r.first / 128; bi * 128 < r.second; ++bi)
                                if (bi >= bj) pl.tiles.push_back(make_int2(bi, bj));
                        }
                    d.seg[t + 1] = (int)pl.tiles.size();
                }
            }
        }
        pn.col_end = (int)pl.tiles.size();
    }
    (void)myp;
    return (int)pl.tiles.size();
}

void dist_free_plan(nes_ctx* c, DistPlan* pl) {
    if (!pl) return;
    for (auto& v : {&pl->ev_arrived, &pl->ev_packed, &pl->ev_colready, &pl->ev_diagdone}) {
        for (cudaEvent_t e : *v)
            if (e) cudaEventDestroy(e);
        v->clear();
    }
    if (pl->ev_start) cudaEventDestroy(pl->ev_start);
    pl->ev_start = nullptr;
    for (auto& e : pl->ev_end) {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
    for (auto& p : pl->d_stage) {
        dev_free(c, p);
        p = nullptr;
    }
}

int dist_broadcast(nes_ctx* c, double* d_buf, size_t count, int root, cudaStream_t stream) {
    NcclApi* api = nccl_api();
    if (!api || !c->nccl_comm) return fail(c, NES_ERR_COMM, "NCCL communicator not initialised");
    ncclResult_t r = api->Broadcast(d_buf, d_buf, count, ncclDouble, root,
                                    static_cast<ncclComm_t>(c->nccl_comm), stream ? stream : c->stream);
    if (r != ncclSuccess)
        return fail(c, NES_ERR_COMM, "ncclBroadcast failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    return 0;
}

int dist_allreduce_int(nes_ctx* c, int* d_buf, size_t count, int op_max_else_min) {
    NcclApi* api = nccl_api();
    if (!api || !c->nccl_comm) return fail(c, NES_ERR_COMM, "NCCL communicator not initialised");
    ncclResult_t r = api->AllReduce(d_buf, d_buf, count, ncclInt32, op_max_else_min ? ncclMax : ncclMin,
                                    static_cast<ncclComm_t>(c->nccl_comm), c->stream);
    if (r != ncclSuccess)
        return fail(c, NES_ERR_COMM, "ncclAllReduce failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    return 0;
}

int dist_allreduce_sum(nes_ctx* c, double* d_buf, size_t count) {
    NcclApi* api = nccl_api();
    if (!api || !c->nccl_comm) return fail(c, NES_ERR_COMM, "NCCL communicator not initialised");
    ncclResult_t r = api->AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum,
                                    static_cast<ncclComm_t>(c->nccl_comm), c->stream);
    if (r != ncclSuccess)
        return fail(c, NES_ERR_COMM, "ncclAllReduce failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    return 0;
}

}  // namespace nes

using namespace nes;

extern "C" {

int nes_comm_unique_id(unsigned char* id128) {
    NcclApi* api = nccl_api();
    if (!api || !id128) return NES_ERR_COMM;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return NES_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return 0;
}

int nes_comm_init(nes_ctx* c, int nranks, int rank, const unsigned char* id128) {
    NES_ENTER(c);
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, NES_ERR_INVALID, "bad rank %d / %d", rank, nranks);
    if (c->nccl_comm) return fail(c, NES_ERR_INVALID, "communicator already initialised");
    c->nranks = nranks;
    c->rank = rank;
    c->grid_p = 1;
    c->grid_q = nranks;
    if (const char* e = getenv("NES_DIST_GRID")) {  // "PxQ", every rank must see the same value
        int gp = 0, gq = 0;
        if (sscanf(e, "%dx%d", &gp, &gq) == 2 && gp >= 1 && gq >= 1 && gp * gq == nranks) {
            c->grid_p = gp;
            c->grid_q = gq;
        } else {
            return fail(c, NES_ERR_INVALID, "NES_DIST_GRID=%s does not describe a grid of %d ranks", e, nranks);
        }
    }
    if (nranks == 1) return 0;
    NcclApi* api = nccl_api();
    if (!api) return fail(c, NES_ERR_COMM, "libnccl.so.2 not found");
    if (!id128) return fail(c, NES_ERR_INVALID, "null unique id");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    // A dmma_nt CTA needs every register of its SM, so each resident NCCL CTA takes one SM away from the
    // trailing updates for as long as a receiver waits for its panel (32 channels by default on NVSwitch:
    // measured 217 -> 204 ms per factorization at m = 32768 on 2 GPUs with 8).  The panels are small against
    // NVLink bandwidth, a few CTAs move them fast enough.  NES_NCCL_CTAS overrides (0 = NCCL's default).
    int max_ctas = 8;
    if (const char* e = getenv("NES_NCCL_CTAS")) max_ctas = atoi(e);
    ncclResult_t r;
    if (api->CommInitRankConfig && max_ctas > 0) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = max_ctas;
        r = api->CommInitRankConfig(&comm, nranks, id, rank, &cfg);
    } else {
        r = api->CommInitRank(&comm, nranks, id, rank);
    }
    if (r != ncclSuccess)
        return fail(c, NES_ERR_COMM, "ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    c->nccl_comm = comm;
    return 0;
}

int nes_comm_finalize(nes_ctx* c) {
    if (!c) return NES_ERR_INVALID;
    if (c->nccl_comm) {
        NcclApi* api = nccl_api();
        if (c->started) cudaStreamSynchronize(c->stream);
        if (api) api->CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
        c->nccl_comm = nullptr;
    }
    c->nranks = 1;
    c->rank = 0;
    c->grid_p = c->grid_q = 1;
    return 0;
}

// Layout the dense factorization uses for an m x m normal matrix on this context's ranks: distribution
// block (= outer panel width) and the P x Q process grid.
int nes_dist_layout(const nes_ctx* c, int m, int* nbo, int* P, int* Q) {
    const int nr = c ? c->nranks : 1;
    if (nbo) *nbo = dense_outer_block(m, nr);
    if (P) *P = c ? c->grid_p : 1;
    if (Q) *Q = c ? c->grid_q : nr;
    return 0;
}

// Process grid of the dense factorization (default 1 x nranks, or NES_DIST_GRID at nes_comm_init): every rank
// must set the same grid, before the factors that use it are analyzed.
int nes_dist_set_grid(nes_ctx* c, int P, int Q) {
    if (!c) return NES_ERR_INVALID;
    if (P < 1 || Q < 1 || P * Q != c->nranks) return fail(c, NES_ERR_INVALID, "grid %d x %d is not %d ranks", P, Q, c->nranks);
    c->grid_p = P;
    c->grid_q = Q;
    return 0;
}

int nes_comm_rank(const nes_ctx* c) { return c ? c->rank : 0; }
int nes_comm_nranks(const nes_ctx* c) { return c ? c->nranks : 1; }

// Host-only planner, exposed so the partition can be checked without a GPU: writes up to `cap`
// (tile row, tile column) pairs of the tiles `rank` owns on a 1 x nranks grid and returns how many there are.
int nes_dist_plan(int m, int nranks, int rank, int* tile_rows, int* tile_cols, int cap) {
    return nes_dist_plan_grid(m, 0, 1, nranks, rank, 0, tile_rows, tile_cols, cap, nullptr, nullptr);
}

// The same for a P x Q grid, distribution block nbo (0 = default) and chunk_rows (0 = default); also reports
// the number of broadcasts of the whole factorization and how many of them this rank is the root of.
int nes_dist_plan_grid(int m, int nbo, int P, int Q, int rank, int chunk_rows, int* tile_rows, int* tile_cols,
                       int cap, int* nmsgs, int* nroot) {
    if (m <= 0 || P < 1 || Q < 1 || rank < 0 || rank >= P * Q) return NES_ERR_INVALID;
    if (nbo <= 0) nbo = dense_outer_block(m, P * Q);
    if (nbo != 128 && nbo != 256 && nbo != 512 && nbo != 1024) return NES_ERR_INVALID;
    if (chunk_rows <= 0) chunk_rows = dist_chunk_rows(m, nbo, P);
    if (chunk_rows % (P * nbo) != 0) return NES_ERR_INVALID;
    DistPlan pl;
    const int n = dist_make_plan(pl, m, nbo, P, Q, rank, chunk_rows);
    for (int i = 0; i < n && i < cap; ++i) {
        if (tile_rows) tile_rows[i] = pl.tiles[i].x;
        if (tile_cols) tile_cols[i] = pl.tiles[i].y;
    }
    int total = 0, mine = 0;
    for (const DistPanel& pn : pl.panels)
        for (const DistMsg& d : pn.msgs) {
            ++total;
            if (d.root == rank) ++mine;
        }
    if (nmsgs) *nmsgs = total;
    if (nroot) *nroot = mine;
    return n;
}

// The message schedule itself (host only): for every broadcast of the factorization, in order, 8 ints
// { panel J, root, has_diag, row_start, nblocks, bh, stride, dep } (dep = index, within panel J-1, of the last
// message the root needs before it may update these rows; -1 = none).  Returns the number of messages; fills up
// to `cap` of them.  tests/dist_worker.py replays the schedule in NumPy over gloo with it.
int nes_dist_plan_msgs(int m, int nbo, int P, int Q, int chunk_rows, int head_blocks, int* out, int cap) {
    if (m <= 0 || P < 1 || Q < 1) return NES_ERR_INVALID;
    if (nbo <= 0) nbo = dense_outer_block(m, P * Q);
    if (chunk_rows <= 0) chunk_rows = dist_chunk_rows(m, nbo, P);
    if (chunk_rows % (P * nbo) != 0) return NES_ERR_INVALID;
    DistPlan pl;
    dist_make_plan(pl, m, nbo, P, Q, 0, chunk_rows, head_blocks);
    int n = 0;
    for (int J = 0; J < pl.nblk; ++J)
        for (const DistMsg& d : pl.panels[J].msgs) {
            if (out && n < cap) {
                int* o = out + 8 * n;
                o[0] = J; o[1] = d.root; o[2] = d.has_diag; o[3] = d.row_start;
                o[4] = d.nblocks; o[5] = d.bh; o[6] = d.stride; o[7] = d.dep;
            }
            ++n;
        }
    return n;
}

}  // extern "C"
