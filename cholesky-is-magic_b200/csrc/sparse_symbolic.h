// Host-side symbolic analysis of A diag(theta) A' for the supernodal multifrontal Cholesky
// (sparse_chol.cu).  Plays cholmod_analyze (sparse-cholesky.lisp:261, 509; affine-scaling.lisp:270-271):
// runs once per sparsity pattern, produces the fill-reducing permutation, the supernode partition with
// its row structures, the assembly tree with its level schedule, the index maps of the numeric phase
// and the counters the reference prints (anz, aatfl, lnz, fl; affine-scaling.lisp:273-279).
// Pure host code: no CUDA calls, usable (and tested) without a GPU.
#pragma once
#include <cstddef>
#include <vector>

namespace nes {

struct Symbolic {
    int m = 0, nsuper = 0, nlevels = 0;
    std::vector<int> perm;          // perm[new] = old
    // supernode s owns columns [first[s], first[s+1]) of P M P'; supernodes are numbered by level of the
    // assembly tree (leaves = level 0), so [lvlptr[l], lvlptr[l+1]) are independent of each other
    std::vector<int> first;         // nsuper + 1
    std::vector<int> nr, ld;        // rows of the supernode's block of L; its leading dimension (16-aligned)
    // Block layout: rows 0..nc-1 are the diagonal block, the nu = nr - nc rows below it start at storage
    // row nb0 = nc rounded up to 32 (the pad rows stay zero), so slabs and TMA boxes of the rows below
    // start 128-byte aligned; ld = nb0 + nu rounded up to 16.
    std::vector<int> nb0;
    std::vector<int> rowptr, rows;  // sorted row lists (permuted numbering); the first nc rows are the columns
    std::vector<long long> off;     // nsuper + 1: offset of the block in the L storage (doubles, 16-aligned)
    long long lsize = 0;
    std::vector<int> sparent;       // assembly-tree parent (-1: root)
    std::vector<int> level, lvlptr;
    std::vector<int> childptr, child;  // children of every supernode, ascending
    // multifrontal maps: below-row i of s (i < nu = nr - nc) is row rel[relptr[s] + i] of the parent's
    // row list; the first cut[s] of them are columns of the parent
    std::vector<int> relptr, rel, cut;
    // tb[tbptr[s] + k] = first below-row i of s whose parent row rel[i] is >= nc_parent + 64 k
    // (k = 0 .. ceil(nu_parent / 64)): the rows of s that land in each 64-row slab of the parent's rows
    // below, precomputed so the extend-add kernels do not search
    std::vector<int> tbptr, tb;
    // update matrices U_s (nu x nu, ld ldu, lower triangle): offsets in a pool whose slots are reused
    // once the parent has consumed them
    std::vector<int> ldu;
    std::vector<long long> uoff;
    long long usize = 0;
    // triangular solves: outgoing segments of s (its below rows grouped by the supernode owning them)
    // and the same segments listed per target, sources ascending
    std::vector<int> segptr, seg_tid, seg_j0, seg_j1;
    std::vector<int> inptr, in_s, in_j0, in_j1;
    // assembly: entry e of tril(P M P') = rows ei[e], ej[e] of A (original numbering) -> edest[e] in L
    std::vector<int> ei, ej;
    std::vector<long long> edest;
    long long anz = 0;
    double aatfl = 0, lnz = 0, fl = 0;
    // multi-GPU: owner[s] = rank whose subtree holds s, or -1 for the top of the tree, which every rank
    // factors redundantly after the update matrices of the subtree roots have been exchanged
    std::vector<int> owner;
    std::vector<double> sflops;     // dense flops of each supernode (potrf + trsm + syrk)
    // Update VECTORS of the multifrontal triangular solves: supernode s keeps nu_s doubles at vptr[s].
    // Subtree roots whose parent is in the replicated top ("exchange roots") come first, grouped by
    // owner, both here and in the update-matrix pool (slots [xu_off[q], xu_off[q+1]) never reused), so
    // a rank publishes everything the top needs from it with one broadcast.
    std::vector<long long> vptr;    // nsuper + 1 entries are not prefix sums: use nu_s for the length
    long long vsize = 0;
    std::vector<long long> xu_off, xv_off;  // nranks + 1 (empty when nranks == 1)
};

struct SymbolicOptions {
    int nranks = 1;
    int nd_leaf = 0;  // nested dissection stops at vertex sets of this size (0 = max(256, m / 128))
};

// 0 on success; negative on an internal inconsistency (message in err)
int symbolic_analyze(int m, int n, const int* colptr, const int* rowidx, const SymbolicOptions& opt,
                     Symbolic* out, char* err, size_t errlen);

}  // namespace nes
