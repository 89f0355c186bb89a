// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Distribution: the normal matrix M and its factor are laid out block-cyclically
// by outer block columns (width NBO) over a 1 x Q process grid -- the 2D block-cyclic scheme with P = 1.
// With NVSwitch every GPU receives a broadcast at full link bandwidth, so nothing is gained by also
// splitting panel rows (P > 1) at 2..8 GPUs, while P = 1 keeps every panel factorization (diagonal block
// + TRSM) local to its owner: the only exchange step is one ncclBroadcast of the finished panel per
// outer block column (<= 134 MB at m = 32768, NBO = 512).  A and the IPM vectors are replicated
// (A is regenerated or uploaded once per rank; all ranks run the O(n) vector kernels redundantly and
// deterministically, so no scalar ever needs to be exchanged).
//
// NCCL is resolved at run time (dlopen) so the library still loads on hosts without it; the unique id
// travels through the caller's launcher (torch.distributed in bench.py / tests).
#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

#include "nes_internal.h"

namespace nes {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
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
        if (v == 128 || v == 256 || v == 512) return v;
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

// Owned tiles (128 x 128) of the lower triangle for `rank`: tile column bj belongs to outer block
// column J = bj*128 / nbo, owned by dist_owner(J).  Ordered by J, then tile row, then tile column, so that
// (a) the tiles updated after panel J are a suffix of the list (from tile_first[J + 1]) and (b) consecutive tiles share rows.
int dist_plan_tiles(int m, int nbo, int nranks, int rank, std::vector<int2>& tiles,
                    std::vector<int>& tile_first) {
    const int tm = (m + 127) / 128;
    const int tpb = nbo / 128;  // tile columns per outer block
    const int nblk = (m + nbo - 1) / nbo;
    tiles.clear();
    tile_first.assign(nblk + 1, 0);
    for (int J = 0; J < nblk; ++J) {
        tile_first[J] = (int)tiles.size();  // first tile with block column >= J (fixed up below)
        if (dist_owner(J, nranks) != rank) continue;
        const int c_lo = J * tpb, c_hi = (c_lo + tpb < tm) ? c_lo + tpb : tm;
        for (int bi = c_lo; bi < tm; ++bi)
            for (int bj = c_lo; bj < c_hi && bj <= bi; ++bj) tiles.push_back(make_int2(bi, bj));
    }
    tile_first[nblk] = (int)tiles.size();
    // tile_first[J] = first owned tile with block column >= J; the update after panel J starts at
    // tile_first[J + 1]
    return (int)tiles.size();
}

int dist_broadcast(nes_ctx* c, double* d_buf, size_t count, int root) {
    NcclApi* api = nccl_api();
    if (!api || !c->nccl_comm) return fail(c, NES_ERR_COMM, "NCCL communicator not initialised");
    ncclResult_t r = api->Broadcast(d_buf, d_buf, count, ncclDouble, root,
                                    static_cast<ncclComm_t>(c->nccl_comm), c->stream);
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
    if (nranks == 1) return 0;
    NcclApi* api = nccl_api();
    if (!api) return fail(c, NES_ERR_COMM, "libnccl.so.2 not found");
    if (!id128) return fail(c, NES_ERR_INVALID, "null unique id");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    ncclResult_t r = api->CommInitRank(&comm, nranks, id, rank);
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
    return 0;
}

// Layout the dense factorization uses for an m x m normal matrix on this context's ranks: distribution
// block (= outer panel width) and the P x Q process grid.
int nes_dist_layout(const nes_ctx* c, int m, int* nbo, int* P, int* Q) {
    const int nr = c ? c->nranks : 1;
    if (nbo) *nbo = dense_outer_block(m, nr);
    if (P) *P = 1;
    if (Q) *Q = nr;
    return 0;
}

int nes_comm_rank(const nes_ctx* c) { return c ? c->rank : 0; }
int nes_comm_nranks(const nes_ctx* c) { return c ? c->nranks : 1; }

// Host-only planner, exposed so the partition can be checked without a GPU: writes up to `cap`
// (tile row, tile column) pairs of the tiles `rank` owns and returns how many there are.
int nes_dist_plan(int m, int nranks, int rank, int* tile_rows, int* tile_cols, int cap) {
    if (m <= 0 || nranks < 1 || rank < 0 || rank >= nranks) return NES_ERR_INVALID;
    std::vector<int2> tiles;
    std::vector<int> first;
    const int n = dist_plan_tiles(m, dense_outer_block(m, nranks), nranks, rank, tiles, first);
    for (int i = 0; i < n && i < cap; ++i) {
        if (tile_rows) tile_rows[i] = tiles[i].x;
        if (tile_cols) tile_cols[i] = tiles[i].y;
    }
    return n;
}

}  // extern "C"
