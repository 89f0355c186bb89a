// Symbolic analysis for the sparse path (host only; see sparse_symbolic.h).
//
//   1. pattern of A A' (the reference hands CHOLMOD the unsymmetric A, stype 0, and CHOLMOD analyzes
//      A A': sparse-cholesky.lisp:408, 509)
//   2. ordering: rows denser than 10 sqrt(m) last; the rest by George's automatic nested dissection
//      (BFS level structure from a pseudo-peripheral vertex, the middle level trimmed to the vertices
//      that touch the far side is the separator, recursively) with reverse Cuthill-McKee inside the
//      leaves.  CHOLMOD would use AMD; the factor and the solution do not depend on the ordering beyond
//      rounding, only lnz / fl do.  Dissection is what gives the GPU independent subtrees to work on
//      at the same time (and the ranks of a multi-GPU run subtrees to own): on a banded pattern a
//      profile ordering leaves a chain of ~m/128 supernodes, each waiting for the previous one.
//   3. elimination tree (Liu, path compression), postorder, column counts (Gilbert-Ng-Peyton skeleton
//      counting) -- near-linear in nnz(A A'), no per-column structures
//   4. supernodes: etree chains, relaxed amalgamation (<= 128 columns, <= 30% explicit zeros)
//   5. assembly tree levels; supernodes renumbered by level (a topological order, so fill is unchanged)
//   6. supernodal row structures, parent-relative index maps, update-matrix pool with slot reuse,
//      solve segments, assembly map, subtree-to-rank mapping
#include "sparse_symbolic.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <atomic>
#include <numeric>
#include <thread>
#ifdef __linux__
#include <sched.h>
#endif

namespace nes {

namespace {

constexpr int SN_MAX_COLS = 128;

// The two adjacency arrays (A A' and its permuted copy, 4e7 ints each at config 4) are written in full
// right after they are sized: resize() must not value-initialise them first (a sequential 160 MB fill).
template <class T>
struct default_init_alloc : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = default_init_alloc<U>;
    };
    template <class U, class... Args>
    void construct(U* p, Args&&... args) {
        if constexpr (sizeof...(Args) == 0)
            ::new (static_cast<void*>(p)) U;
        else
            ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
    }
};
using adj_vec = std::vector<int, default_init_alloc<int>>;

int host_threads() {
    static int n = 0;
    if (n == 0) {
        n = (int)std::thread::hardware_concurrency();
#ifdef __linux__
        {   // a cpuset smaller than the machine (containers): more threads than CPUs only slow the barriers down
            cpu_set_t set;
            CPU_ZERO(&set);
            if (sched_getaffinity(0, sizeof(set), &set) == 0) {
                const int allowed = CPU_COUNT(&set);
                if (allowed > 0 && allowed < n) n = allowed;
            }
        }
#endif
        if (const char* e = getenv("NES_HOST_THREADS")) n = atoi(e);
        n = n < 1 ? 1 : (n > 32 ? 32 : n);
    }
    return n;
}

// f(chunk, begin, end) on contiguous chunks of [0, n); chunk ids are 0 .. nchunks-1 in index order
template <typename F>
void parallel_chunks(int n, int nchunks, F f) {
    if (nchunks <= 1 || n < 4096) {
        f(0, 0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int k = 0; k < nchunks; ++k) {
        const int b = (int)((long long)n * k / nchunks), e = (int)((long long)n * (k + 1) / nchunks);
        th.emplace_back([=]() { f(k, b, e); });
    }
    for (auto& t : th) t.join();
}

void csc_to_csr(int m, int n, const int* cp, const int* ri, std::vector<int>& rp, std::vector<int>& cj) {
    rp.assign(m + 1, 0);
    cj.resize(cp[n]);
    for (int k = 0; k < cp[n]; ++k) rp[ri[k] + 1]++;
    for (int i = 0; i < m; ++i) rp[i + 1] += rp[i];
    std::vector<int> next(rp.begin(), rp.end() - 1);
    for (int j = 0; j < n; ++j)
        for (int k = cp[j]; k < cp[j + 1]; ++k) cj[next[ri[k]]++] = j;
}

// full symmetric adjacency of A A' (without the diagonal), sorted per row; rows are independent, so
// chunks of rows are built by host threads and concatenated
void aat_pattern(int m, int n, const int* cp, const int* ri, const std::vector<int>& rp,
                 const std::vector<int>& cj, std::vector<int>& ap, adj_vec& ai, double& aatfl) {
    ap.assign(m + 1, 0);
    aatfl = 0;
    for (int j = 0; j < n; ++j) {
        const double c = cp[j + 1] - cp[j];
        aatfl += c * c;
    }
    const int nch = (m < 4096) ? 1 : host_threads();
    std::vector<std::vector<int>> part(nch);
    parallel_chunks(m, nch, [&](int k, int b, int e) {
        std::vector<int> mark(m, -1);
        std::vector<int>& out = part[k];
        for (int i = b; i < e; ++i) {
            const size_t start = out.size();
            mark[i] = i;
            int lo = m, hi = -1;
            for (int q = rp[i]; q < rp[i + 1]; ++q) {
                const int kk = cj[q];
                for (int t = cp[kk]; t < cp[kk + 1]; ++t) {
                    const int r = ri[t];
                    if (mark[r] != i) {
                        mark[r] = i;
                        out.push_back(r);
                        lo = std::min(lo, r);
                        hi = std::max(hi, r);
                    }
                }
            }
            const int cnt = (int)(out.size() - start);
            if (cnt > 32 && hi - lo + 1 <= 4 * cnt) {
                // the neighbours fill a short index range (banded / clustered rows): reading the marks of
                // that range in order is cheaper than sorting
                size_t w = start;
                for (int r = lo; r <= hi; ++r)
                    if (mark[r] == i && r != i) out[w++] = r;
            } else {
                std::sort(out.begin() + start, out.end());
            }
            ap[i + 1] = cnt;
        }
    });
    for (int i = 0; i < m; ++i) ap[i + 1] += ap[i];
    ai.resize(ap[m]);
    parallel_chunks(m, nch, [&](int k, int b, int) {  // same chunk boundaries: chunk k starts at row b
        std::copy(part[k].begin(), part[k].end(), ai.begin() + ap[b]);
    });
}

// ---- ordering ---------------------------------------------------------------------------------------
struct Orderer {
    int m;
    const std::vector<int>& ap;
    const adj_vec& ai;
    std::vector<int> part;   // label of the vertex set a vertex currently belongs to (-1: removed)
    std::vector<int> deg;
    std::vector<int> lev;    // scratch: BFS level (sets being dissected concurrently are disjoint)
    std::vector<int> claim;  // scratch of bfs_parallel (INT_MAX between searches)
    std::atomic<int> next_label{1};
    int leaf;

    Orderer(int m_, const std::vector<int>& ap_, const adj_vec& ai_, int leaf_)
        : m(m_), ap(ap_), ai(ai_), part(m_, 0), deg(m_), lev(m_, -1), claim(m_, INT_MAX), leaf(leaf_) {
        for (int i = 0; i < m; ++i) deg[i] = ap[i + 1] - ap[i];
    }

    // BFS inside label `lab` from `start`; fills `order` (visit order) and lev[]; returns the number of levels
    int bfs(int start, int lab, std::vector<int>& order) {
        order.clear();
        order.push_back(start);
        lev[start] = 0;
        size_t head = 0;
        int last = 0;
        while (head < order.size()) {
            const int v = order[head++];
            last = lev[v];
            for (int q = ap[v]; q < ap[v + 1]; ++q) {
                const int u = ai[q];
                if (part[u] == lab && lev[u] < 0) {
                    lev[u] = last + 1;
                    order.push_back(u);
                }
            }
        }
        return last + 1;
    }
    void clear_lev(const std::vector<int>& order) {
        for (int v : order) lev[v] = -1;
    }

    // The same search on `nth` host threads, level by level, with EXACTLY the visit order of bfs(): a vertex
    // belongs to the first frontier vertex (in queue order) that reaches it and is appended at its place in
    // that vertex's adjacency list.  Pass 1 lets every frontier vertex claim its unvisited neighbours
    // (minimum of the queue positions), pass 2 emits the neighbours a vertex won, in adjacency order, into
    // per-thread lists that are concatenated in thread (= queue) order.  Two passes over the edges instead
    // of one, but on all threads: the top-level searches of the dissection walk the whole graph (4e7 edges
    // at config 4) and were the longest sequential piece of the analysis.
    int bfs_parallel(int start, int lab, std::vector<int>& order, int nth) {
        order.clear();
        order.push_back(start);
        lev[start] = 0;
        struct Shared {
            std::atomic<int> arrived{0}, phase{0};
            size_t fb = 0, fe = 1;   // current frontier = order[fb, fe)
            int level = 0;
            bool done = false;
        } sh;
        std::vector<std::vector<int>> found(nth);
        auto barrier = [&](int& my_phase) {
            ++my_phase;
            if (sh.arrived.fetch_add(1, std::memory_order_acq_rel) == nth - 1) {
                sh.arrived.store(0, std::memory_order_relaxed);
                sh.phase.store(my_phase, std::memory_order_release);
            } else {
                int spins = 0;
                while (sh.phase.load(std::memory_order_acquire) != my_phase)
                    if (++spins > 2000) std::this_thread::yield();
            }
        };
        auto worker = [&](int t) {
            int my_phase = 0;
            for (;;) {
                const size_t fb = sh.fb, fe = sh.fe, nf = fe - fb;
                const size_t a = fb + nf * t / nth, b = fb + nf * (t + 1) / nth;
                for (size_t p = a; p < b; ++p) {   // pass 1: claim
                    const int v = order[p];
                    for (int q = ap[v]; q < ap[v + 1]; ++q) {
                        const int u = ai[q];
                        if (part[u] != lab || lev[u] >= 0) continue;
                        int cur = __atomic_load_n(&claim[u], __ATOMIC_RELAXED);
                        while ((int)p < cur &&
                               !__atomic_compare_exchange_n(&claim[u], &cur, (int)p, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                        }
                    }
                }
                barrier(my_phase);
                std::vector<int>& mine = found[t];
                mine.clear();
                for (size_t p = a; p < b; ++p) {   // pass 2: emit what this vertex won, in adjacency order
                    const int v = order[p];
                    for (int q = ap[v]; q < ap[v + 1]; ++q) {
                        const int u = ai[q];
                        if (part[u] == lab && lev[u] < 0 && claim[u] == (int)p) mine.push_back(u);
                    }
                }
                barrier(my_phase);
                if (t == 0) {                      // next frontier, in queue order
                    for (int k = 0; k < nth; ++k)
                        for (int u : found[k]) {
                            lev[u] = sh.level + 1;
                            claim[u] = INT_MAX;
                            order.push_back(u);
                        }
                    sh.fb = fe;
                    sh.fe = order.size();
                    if (sh.fe > sh.fb) ++sh.level;
                    sh.done = sh.fe == sh.fb;
                }
                barrier(my_phase);
                if (sh.done) return;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nth; ++t) th.emplace_back(worker, t);
        worker(0);
        for (auto& x : th) x.join();
        return sh.level + 1;
    }
    // A branch of the dissection that may still fork `par_depth` more times owns 2^par_depth host threads:
    // its searches use them while the vertex set is large (the top-level ones walk the whole graph).
    int bfs_top(int start, int lab, std::vector<int>& order, int par_depth, size_t nverts) {
        const int nth = std::min(host_threads(), 1 << std::min(par_depth, 10));
        if (nth > 1 && nverts >= 20000) return bfs_parallel(start, lab, order, nth);
        return bfs(start, lab, order);
    }

    // reverse Cuthill-McKee of the vertices carrying label `lab` (they are relabelled -1 = done),
    // appended to `out`
    void rcm(const std::vector<int>& verts, int lab, std::vector<int>& out) {
        std::vector<int> byDeg(verts);
        std::stable_sort(byDeg.begin(), byDeg.end(), [&](int a, int b) { return deg[a] < deg[b]; });
        std::vector<int> order, nbr;
        order.reserve(verts.size());
        for (int s : byDeg) {
            if (part[s] != lab) continue;
            size_t head = order.size();
            order.push_back(s);
            part[s] = -1;
            while (head < order.size()) {
                const int v = order[head++];
                nbr.clear();
                for (int q = ap[v]; q < ap[v + 1]; ++q) {
                    const int u = ai[q];
                    if (part[u] == lab) {
                        part[u] = -1;
                        nbr.push_back(u);
                    }
                }
                std::sort(nbr.begin(), nbr.end(),
                          [&](int a, int b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
                order.insert(order.end(), nbr.begin(), nbr.end());
            }
        }
        out.insert(out.end(), order.rbegin(), order.rend());
    }

    // elimination order of `verts` (all labelled `lab`) appended to `out`.  `connected`: the set is known to
    // be one component (the near side of a level-structure cut always is).  The two sides of a cut are
    // independent, so the first `par_depth` levels of the recursion run them on two host threads; the
    // result does not depend on the number of threads.
    void dissect(std::vector<int> verts, int lab, int depth, bool connected, int par_depth, std::vector<int>& out) {
        if ((int)verts.size() <= leaf || depth > 48) {
            rcm(verts, lab, out);
            return;
        }
        // pseudo-peripheral start: min degree, then the far end of its level structure
        int start = verts[0];
        for (int v : verts)
            if (deg[v] < deg[start]) start = v;
        std::vector<int> order;
        int h = bfs_top(start, lab, order, par_depth, verts.size());
        if (!connected && order.size() != verts.size()) {
            // more than one component: they are independent subtrees of the elimination forest
            // (enumerated in the order of `verts`, as before the first search doubled as the connectivity test)
            clear_lev(order);
            std::vector<std::vector<int>> comps;
            for (int s : verts) {
                if (lev[s] >= 0) continue;
                bfs(s, lab, order);
                comps.push_back(order);
            }
            for (auto& cvec : comps) clear_lev(cvec);
            for (auto& cvec : comps) {
                const int l2 = next_label++;
                for (int v : cvec) part[v] = l2;
                dissect(std::move(cvec), l2, depth + 1, true, 0, out);
            }
            return;
        }
        {
            const int far = order.back();
            clear_lev(order);
            h = bfs_top(far, lab, order, par_depth, verts.size());
        }
        const int nv = (int)verts.size();
        if (h < 3) {
            clear_lev(order);
            rcm(verts, lab, out);
            return;
        }
        std::vector<int> lsize(h, 0);
        for (int v : order) lsize[lev[v]]++;
        // separator level: smallest level that leaves at least 30% of the vertices on each side; if no
        // level does, the one where the running count crosses the middle
        int k = -1;
        {
            int before = lsize[0];
            for (int l = 1; l + 1 < h; ++l) {
                const int after = nv - before - lsize[l];
                if (before >= 0.3 * nv && after >= 0.3 * nv && (k < 0 || lsize[l] < lsize[k])) k = l;
                before += lsize[l];
            }
            if (k < 0) {
                before = lsize[0];
                for (int l = 1; l + 1 < h; ++l) {
                    if (before + lsize[l] >= nv / 2) {
                        k = l;
                        break;
                    }
                    before += lsize[l];
                }
                if (k < 0) k = h / 2;
            }
        }
        if (lsize[k] * 3 > nv) {  // no useful separator (close to a clique)
            clear_lev(order);
            rcm(verts, lab, out);
            return;
        }
        const int la = next_label++, lb = next_label++, ls = next_label++;
        std::vector<int> A, B, S;
        for (int v : order) {
            if (lev[v] < k) {
                A.push_back(v);
            } else if (lev[v] > k) {
                B.push_back(v);
            } else {
                bool touches = false;
                for (int q = ap[v]; q < ap[v + 1] && !touches; ++q) {
                    const int u = ai[q];
                    touches = part[u] == lab && lev[u] == k + 1;
                }
                (touches ? S : A).push_back(v);
            }
        }
        clear_lev(order);
        for (int v : A) part[v] = la;
        for (int v : B) part[v] = lb;
        for (int v : S) part[v] = ls;
        order.clear();
        order.shrink_to_fit();
        verts.clear();
        verts.shrink_to_fit();
        if (par_depth > 0 && A.size() > 4096 && B.size() > 4096) {
            std::vector<int> outA, outB;
            std::thread ta([&]() { dissect(std::move(A), la, depth + 1, true, par_depth - 1, outA); });
            dissect(std::move(B), lb, depth + 1, false, par_depth - 1, outB);
            ta.join();
            out.insert(out.end(), outA.begin(), outA.end());
            out.insert(out.end(), outB.begin(), outB.end());
        } else {
            dissect(std::move(A), la, depth + 1, true, 0, out);
            dissect(std::move(B), lb, depth + 1, false, 0, out);
        }
        rcm(S, ls, out);
    }
};

void order_nd(int m, const std::vector<int>& ap, const adj_vec& ai, int leaf, std::vector<int>& perm) {
    Orderer o(m, ap, ai, leaf);
    const double thresh = std::max(16.0, 10.0 * std::sqrt((double)m));
    std::vector<int> sparse_rows, dense_rows;
    for (int i = 0; i < m; ++i) {
        if (o.deg[i] > thresh) {
            dense_rows.push_back(i);
            o.part[i] = -1;
        } else {
            sparse_rows.push_back(i);
        }
    }
    int par_depth = 0;
    for (int t = host_threads(); t > 1; t >>= 1) ++par_depth;
    perm.clear();
    perm.reserve(m);
    o.dissect(std::move(sparse_rows), 0, 0, false, par_depth, perm);
    perm.insert(perm.end(), dense_rows.begin(), dense_rows.end());
}

// adjacency in the permuted numbering (unsorted)
void permute_graph(int m, const std::vector<int>& ap, const adj_vec& ai, const std::vector<int>& perm,
                   const std::vector<int>& iperm, std::vector<int>& pp, adj_vec& pi) {
    pp.assign(m + 1, 0);
    pi.resize(ai.size());
    for (int j = 0; j < m; ++j) pp[j + 1] = pp[j] + (ap[perm[j] + 1] - ap[perm[j]]);
    parallel_chunks(m, host_threads(), [&](int, int b, int e) {
        for (int j = b; j < e; ++j) {
            int w = pp[j];
            for (int q = ap[perm[j]]; q < ap[perm[j] + 1]; ++q) pi[w++] = iperm[ai[q]];
        }
    });
}

void etree(int m, const std::vector<int>& pp, const adj_vec& pi, std::vector<int>& parent) {
    parent.assign(m, -1);
    std::vector<int> anc(m, -1);
    for (int j = 0; j < m; ++j) {
        for (int q = pp[j]; q < pp[j + 1]; ++q) {
            int r = pi[q];
            if (r >= j) continue;
            while (anc[r] != -1 && anc[r] != j) {
                const int nx = anc[r];
                anc[r] = j;
                r = nx;
            }
            if (anc[r] == -1) {
                anc[r] = j;
                parent[r] = j;
            }
        }
    }
}

// The same tree straight from A (CSR: rp, cj), without walking the 4e7 entries of the pattern of A A':
// the rows of A that share column c form a clique of A A', and for the elimination tree a clique is as
// good as the chain of its rows in elimination order, so linking every row to the previous row of each of
// its columns (prev[c]) gives the tree of A A' in O(nnz(A) alpha) (the "column elimination tree" of A').
void etree_of_aat(int m, int n, const std::vector<int>& rp, const std::vector<int>& cj, const std::vector<int>& perm,
                  std::vector<int>& parent) {
    parent.assign(m, -1);
    std::vector<int> anc(m, -1), prev(n, -1);
    for (int k = 0; k < m; ++k) {
        const int row = perm[k];
        for (int q = rp[row]; q < rp[row + 1]; ++q) {
            const int c = cj[q];
            int i = prev[c];
            while (i != -1 && i < k) {
                const int nx = anc[i];
                anc[i] = k;
                if (nx == -1) parent[i] = k;
                i = nx;
            }
            prev[c] = k;
        }
    }
}

// post[k] = k-th vertex of a depth-first postorder (children visited in ascending order)
void postorder(int m, const std::vector<int>& parent, std::vector<int>& post) {
    std::vector<int> head(m, -1), next(m, -1);
    for (int j = m - 1; j >= 0; --j) {
        if (parent[j] < 0) continue;
        next[j] = head[parent[j]];
        head[parent[j]] = j;
    }
    post.clear();
    post.reserve(m);
    std::vector<int> stack;
    for (int r = 0; r < m; ++r) {
        if (parent[r] >= 0) continue;
        stack.push_back(r);
        while (!stack.empty()) {
            const int v = stack.back();
            const int ch = head[v];
            if (ch < 0) {
                post.push_back(v);
                stack.pop_back();
            } else {
                head[v] = next[ch];
                stack.push_back(ch);
            }
        }
    }
}

// Column counts of L for a POSTORDERED tree (vertex k is the k-th in postorder): skeleton counting.
// For every row subtree T_i (the vertices j < i with L_ij != 0) only its leaves are new entries of
// their columns; the overlap of consecutive leaves is removed at their least common ancestor, found
// with a path-compressed disjoint-set forest.
void column_counts(int m, const std::vector<int>& pp, const adj_vec& pi, const std::vector<int>& parent,
                   std::vector<int>& cc) {
    std::vector<int> firstdesc(m, -1), maxfirst(m, -1), prevleaf(m, -1), anc(m), delta(m, 0);
    for (int k = 0; k < m; ++k) {
        delta[k] = (firstdesc[k] == -1) ? 1 : 0;  // leaves of the etree start with their diagonal
        for (int j = k; j != -1 && firstdesc[j] == -1; j = parent[j]) firstdesc[j] = k;
    }
    for (int i = 0; i < m; ++i) anc[i] = i;
    for (int j = 0; j < m; ++j) {
        if (parent[j] != -1) delta[parent[j]]--;  // j is counted in its own column, not its parent's
        for (int q = pp[j]; q < pp[j + 1]; ++q) {
            const int i = pi[q];
            if (i <= j || firstdesc[j] <= maxfirst[i]) continue;  // j is not a leaf of T_i
            maxfirst[i] = firstdesc[j];
            const int jprev = prevleaf[i];
            prevleaf[i] = j;
            delta[j]++;
            if (jprev != -1) {  // subsequent leaf: remove the shared path above lca(jprev, j)
                int r = jprev;
                while (r != anc[r]) r = anc[r];
                for (int s = jprev; s != r;) {
                    const int nx = anc[s];
                    anc[s] = r;
                    s = nx;
                }
                delta[r]--;
            }
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    cc = delta;
    for (int j = 0; j < m; ++j)
        if (parent[j] != -1) cc[parent[j]] += cc[j];
}

// The same counts from A itself (CSC: cp, ri), without the pattern of A A': every column c of A is a clique
// of A A'; it is handed to the first of its rows in the (postordered) elimination order, and when that row
// j comes up the rows i of the clique are tested as "j is a leaf of the row subtree of i" exactly as above
// (Gilbert-Ng-Peyton for A'A, the `ata` variant of CSparse's cs_counts).  O(nnz(A) alpha) instead of O(nnz(A A')).
void column_counts_of_aat(int m, int n, const int* cp, const int* ri, const std::vector<int>& iperm,
                          const std::vector<int>& parent, std::vector<int>& cc) {
    std::vector<int> firstdesc(m, -1), maxfirst(m, -1), prevleaf(m, -1), anc(m), delta(m, 0);
    for (int k = 0; k < m; ++k) {
        delta[k] = (firstdesc[k] == -1) ? 1 : 0;
        for (int j = k; j != -1 && firstdesc[j] == -1; j = parent[j]) firstdesc[j] = k;
    }
    std::vector<int> head(m + 1, -1), next(n, -1);  // columns of A listed under their first row
    for (int c = n - 1; c >= 0; --c) {
        int k = m;
        for (int t = cp[c]; t < cp[c + 1]; ++t) k = std::min(k, iperm[ri[t]]);
        next[c] = head[k];
        head[k] = c;
    }
    for (int i = 0; i < m; ++i) anc[i] = i;
    for (int j = 0; j < m; ++j) {
        if (parent[j] != -1) delta[parent[j]]--;
        for (int c = head[j]; c != -1; c = next[c]) {
            for (int t = cp[c]; t < cp[c + 1]; ++t) {
                const int i = iperm[ri[t]];
                if (i <= j || firstdesc[j] <= maxfirst[i]) continue;
                maxfirst[i] = firstdesc[j];
                const int jprev = prevleaf[i];
                prevleaf[i] = j;
                delta[j]++;
                if (jprev != -1) {
                    int r = jprev;
                    while (r != anc[r]) r = anc[r];
                    for (int s = jprev; s != r;) {
                        const int nx = anc[s];
                        anc[s] = r;
                        s = nx;
                    }
                    delta[r]--;
                }
            }
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    cc = delta;
    for (int j = 0; j < m; ++j)
        if (parent[j] != -1) cc[parent[j]] += cc[j];
}

int set_err(char* err, size_t n, const char* msg, int a = 0, int b = 0) {
    if (err && n) snprintf(err, n, msg, a, b);
    return -1;
}

}  // namespace

int symbolic_analyze(int m, int n, const int* cp, const int* ri, const SymbolicOptions& opt, Symbolic* S,
                     char* err, size_t errlen) {
    const bool timing = getenv("NES_SYMBOLIC_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[symbolic] %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    *S = Symbolic();
    S->m = m;
    std::vector<int> rp, cj, ap;
    adj_vec ai;
    csc_to_csr(m, n, cp, ri, rp, cj);
    aat_pattern(m, n, cp, ri, rp, cj, ap, ai, S->aatfl);
    S->anz = (long long)m + (long long)ai.size() / 2;
    lap("pattern of A A'");

    // ---- ordering, etree, postorder, column counts ---------------------------------------------------
    const int leaf = opt.nd_leaf > 0 ? opt.nd_leaf : std::max(256, m / 128);
    std::vector<int> perm, iperm(m), pp, parent, cc;
    adj_vec pi;
    order_nd(m, ap, ai, leaf, perm);
    lap("nested dissection ordering");
    if ((int)perm.size() != m) return set_err(err, errlen, "ordering lost vertices (%d of %d)", (int)perm.size(), m);
    for (int i = 0; i < m; ++i) iperm[perm[i]] = i;
    etree_of_aat(m, n, rp, cj, perm, parent);
    if (getenv("NES_SYMBOLIC_CHECK")) {  // the tree from the full pattern of A A' (what this replaced) must be the same
        std::vector<int> parent_ref;
        permute_graph(m, ap, ai, perm, iperm, pp, pi);
        etree(m, pp, pi, parent_ref);
        if (parent_ref != parent) return set_err(err, errlen, "elimination tree from A differs from the tree of the pattern of A A'");
    }
    {
        std::vector<int> post;
        postorder(m, parent, post);
        std::vector<int> perm2(m);
        for (int k = 0; k < m; ++k) perm2[k] = perm[post[k]];
        perm.swap(perm2);
        for (int i = 0; i < m; ++i) iperm[perm[i]] = i;
        std::vector<int> ipost(m), parent2(m);  // same tree, relabelled: vertex k is now k-th in postorder
        for (int k = 0; k < m; ++k) ipost[post[k]] = k;
        for (int k = 0; k < m; ++k) parent2[k] = parent[post[k]] < 0 ? -1 : ipost[parent[post[k]]];
        parent.swap(parent2);
    }
    lap("etree + postorder");
    column_counts_of_aat(m, n, cp, ri, iperm, parent, cc);
    if (getenv("NES_SYMBOLIC_CHECK")) {  // and the counts from the full pattern
        std::vector<int> cc_ref;
        permute_graph(m, ap, ai, perm, iperm, pp, pi);
        column_counts(m, pp, pi, parent, cc_ref);
        if (cc_ref != cc) return set_err(err, errlen, "column counts from A differ from the counts of the pattern of A A'");
    }
    lap("column counts");
    for (int j = 0; j < m; ++j) {
        S->lnz += cc[j];
        S->fl += (double)cc[j] * cc[j];
    }

    // ---- supernodes: chains parent(j) = j+1, relaxed amalgamation --------------------------------------
    std::vector<int> sfirst;  // in postorder numbering
    sfirst.push_back(0);
    double relax = 0.30;  // explicit zeros tolerated in a relaxed supernode (NES_RELAX overrides: experiments)
    if (const char* e = getenv("NES_RELAX")) relax = atof(e);
    {
        int f = 0;
        for (int j = 0; j < m; ++j) {
            bool merge = false;
            if (j + 1 < m && parent[j] == j + 1) {
                const int width = j + 2 - f;
                const long long nrows = width + (long long)(cc[j + 1] - 1);
                long long z = 0;
                for (int q = f; q <= j + 1; ++q) z += (nrows - (q - f)) - cc[q];
                const long long total = nrows * width - (long long)width * (width - 1) / 2;
                if (width <= SN_MAX_COLS && (width <= 4 || z <= relax * total)) merge = true;
            }
            if (!merge) {
                sfirst.push_back(j + 1);
                f = j + 1;
            }
        }
    }
    const int ns = (int)sfirst.size() - 1;
    S->nsuper = ns;
    // assembly tree and levels (postorder numbering: children have smaller numbers)
    std::vector<int> col2sn(m), spar(ns, -1), lvl(ns, 0);
    for (int s = 0; s < ns; ++s)
        for (int j = sfirst[s]; j < sfirst[s + 1]; ++j) col2sn[j] = s;
    for (int s = 0; s < ns; ++s) {
        const int last = sfirst[s + 1] - 1;
        if (parent[last] >= 0) {
            spar[s] = col2sn[parent[last]];
            lvl[spar[s]] = std::max(lvl[spar[s]], lvl[s] + 1);
        }
    }
    // ---- renumber supernodes by (level, postorder index) --------------------------------------------------
    std::vector<int> sorder(ns), snew(ns);
    std::iota(sorder.begin(), sorder.end(), 0);
    std::stable_sort(sorder.begin(), sorder.end(), [&](int a, int b) { return lvl[a] < lvl[b]; });
    for (int k = 0; k < ns; ++k) snew[sorder[k]] = k;
    S->nlevels = ns ? lvl[sorder[ns - 1]] + 1 : 0;
    S->first.assign(ns + 1, 0);
    S->level.resize(ns);
    S->sparent.resize(ns);
    S->lvlptr.assign(S->nlevels + 1, 0);
    {
        std::vector<int> perm2(m);
        int w = 0;
        for (int k = 0; k < ns; ++k) {
            const int s = sorder[k];
            S->first[k] = w;
            for (int j = sfirst[s]; j < sfirst[s + 1]; ++j) perm2[w++] = perm[j];
            S->level[k] = lvl[s];
            S->sparent[k] = spar[s] < 0 ? -1 : snew[spar[s]];
            S->lvlptr[lvl[s] + 1]++;
        }
        S->first[ns] = w;
        for (int l = 0; l < S->nlevels; ++l) S->lvlptr[l + 1] += S->lvlptr[l];
        perm.swap(perm2);
    }
    for (int i = 0; i < m; ++i) iperm[perm[i]] = i;
    S->perm = perm;
    const std::vector<int>& first = S->first;
    for (int s = 0; s < ns; ++s)
        for (int j = first[s]; j < first[s + 1]; ++j) col2sn[j] = s;
    // last-column counts in the new numbering: cc was indexed by postorder column; column j (postorder) of
    // supernode s sits at first[snew[s]] + (j - sfirst[s])
    std::vector<int> cnew(m);
    for (int s = 0; s < ns; ++s)
        for (int j = sfirst[s]; j < sfirst[s + 1]; ++j) cnew[first[snew[s]] + (j - sfirst[s])] = cc[j];

    // children lists
    S->childptr.assign(ns + 1, 0);
    for (int s = 0; s < ns; ++s)
        if (S->sparent[s] >= 0) S->childptr[S->sparent[s] + 1]++;
    for (int s = 0; s < ns; ++s) S->childptr[s + 1] += S->childptr[s];
    S->child.resize(S->childptr[ns]);
    {
        std::vector<int> next(S->childptr.begin(), S->childptr.end() - 1);
        for (int s = 0; s < ns; ++s)
            if (S->sparent[s] >= 0) S->child[next[S->sparent[s]]++] = s;
    }

    lap("supernodes + levels");
    // ---- supernodal row structures -------------------------------------------------------------------------
    // rows(s) = columns of s, then { i > last(s) : i adjacent to a column of s, or i below a child of s }
    S->nr.resize(ns);
    S->ld.resize(ns);
    S->nb0.resize(ns);
    S->rowptr.assign(ns + 1, 0);
    S->off.assign(ns + 1, 0);
    {
        // The supernodes of a level only read the structures of lower levels: levels with many supernodes
        // are split over host threads (one mark array each), the results are appended in supernode order.
        const int nth = host_threads();
        std::vector<std::vector<int>> marks(nth);
        std::vector<std::vector<int>> below(ns);  // sorted rows below the columns of every supernode
        auto structure_of = [&](int k, int sa, int sb) {
            std::vector<int>& mark = marks[k];
            if (mark.empty()) mark.assign(m, -1);
            for (int s = sa; s < sb; ++s) {
                const int f = first[s], l = first[s + 1] - 1;
                std::vector<int>& tmp = below[s];
                for (int j = f; j <= l; ++j) {
                    const int oj = perm[j];
                    for (int q = ap[oj]; q < ap[oj + 1]; ++q) {
                        const int i = iperm[ai[q]];
                        if (i > l && mark[i] != s) {
                            mark[i] = s;
                            tmp.push_back(i);
                        }
                    }
                }
                for (int q = S->childptr[s]; q < S->childptr[s + 1]; ++q) {
                    const int c = S->child[q];
                    const int cnc = first[c + 1] - first[c];
                    for (int p = S->rowptr[c] + cnc; p < S->rowptr[c + 1]; ++p) {
                        const int i = S->rows[p];
                        if (i > l && mark[i] != s) {
                            mark[i] = s;
                            tmp.push_back(i);
                        }
                    }
                }
                std::sort(tmp.begin(), tmp.end());
            }
        };
        for (int lv = 0; lv < S->nlevels; ++lv) {
            const int sa = S->lvlptr[lv], sb = S->lvlptr[lv + 1];
            if (nth > 1 && sb - sa >= 4 * nth) {
                std::vector<std::thread> th;
                for (int k = 0; k < nth; ++k) {
                    const int b = sa + (int)((long long)(sb - sa) * k / nth), e = sa + (int)((long long)(sb - sa) * (k + 1) / nth);
                    th.emplace_back([&structure_of, k, b, e]() { structure_of(k, b, e); });
                }
                for (auto& t : th) t.join();
            } else {
                structure_of(0, sa, sb);
            }
            for (int s = sa; s < sb; ++s) {
                const int f = first[s], l = first[s + 1] - 1, nc = l - f + 1;
                std::vector<int>& tmp = below[s];
                if ((int)tmp.size() != cnew[l] - 1)
                    return set_err(err, errlen, "symbolic analysis: structure of supernode %d disagrees with its column count",
                                   s, 0);
                for (int j = f; j <= l; ++j) S->rows.push_back(j);
                S->rows.insert(S->rows.end(), tmp.begin(), tmp.end());
                S->nr[s] = nc + (int)tmp.size();
                S->nb0[s] = (nc + 31) & ~31;
                S->ld[s] = S->nb0[s] + (((int)tmp.size() + 15) & ~15);
                S->rowptr[s + 1] = (int)S->rows.size();
                S->off[s + 1] = S->off[s] + (long long)S->ld[s] * nc;
                std::vector<int>().swap(tmp);
            }
        }
    }
    S->lsize = S->off[ns];

    lap("supernodal structures");
    // ---- parent-relative maps -----------------------------------------------------------------------------------
    S->relptr.assign(ns + 1, 0);
    S->cut.assign(ns, 0);
    S->sflops.assign(ns, 0.0);
    for (int s = 0; s < ns; ++s) {
        const int nc = first[s + 1] - first[s], nu = S->nr[s] - nc;
        S->relptr[s + 1] = S->relptr[s] + nu;
        const double r = S->nr[s];
        S->sflops[s] = (double)nc * r * r;  // ~ sum over its columns of (rows below)^2, upper bound
    }
    S->rel.resize(S->relptr[ns]);
    for (int s = 0; s < ns; ++s) {
        const int p = S->sparent[s];
        const int nc = first[s + 1] - first[s], nu = S->nr[s] - nc;
        if (nu == 0) continue;
        if (p < 0) return set_err(err, errlen, "symbolic analysis: supernode %d has rows below but no parent", s, 0);
        const int* R = S->rows.data() + S->rowptr[s] + nc;
        const int* Rp = S->rows.data() + S->rowptr[p];
        const int nrp = S->nr[p], plast = first[p + 1] - 1;
        int q = 0, cut = 0;
        for (int i = 0; i < nu; ++i) {
            while (q < nrp && Rp[q] < R[i]) ++q;
            if (q >= nrp || Rp[q] != R[i])
                return set_err(err, errlen, "symbolic analysis: row %d of supernode %d missing in its parent", R[i], s);
            S->rel[S->relptr[s] + i] = q;
            if (R[i] <= plast) ++cut;
        }
        S->cut[s] = cut;
    }
    S->tbptr.assign(ns + 1, 0);
    for (int s = 0; s < ns; ++s) {
        const int p = S->sparent[s];
        int cnt = 0;
        if (p >= 0 && S->nr[s] > first[s + 1] - first[s]) {
            const int nup = S->nr[p] - (first[p + 1] - first[p]);
            cnt = (nup + 63) / 64 + 1;
        }
        S->tbptr[s + 1] = S->tbptr[s] + cnt;
    }
    S->tb.resize(S->tbptr[ns]);
    for (int s = 0; s < ns; ++s) {
        const int cnt = S->tbptr[s + 1] - S->tbptr[s];
        if (cnt == 0) continue;
        const int p = S->sparent[s];
        const int ncp = first[p + 1] - first[p], nu = S->nr[s] - (first[s + 1] - first[s]);
        const int* rl = S->rel.data() + S->relptr[s];
        int i = S->cut[s];
        for (int k = 0; k < cnt; ++k) {
            while (i < nu && rl[i] < ncp + 64 * k) ++i;
            S->tb[S->tbptr[s] + k] = i;
        }
    }

    // ---- subtree-to-rank mapping ----------------------------------------------------------------------------
    S->owner.assign(ns, 0);
    if (opt.nranks > 1 && ns > 0) {
        const int Q = opt.nranks;
        std::vector<double> w(S->sflops);
        for (int s = 0; s < ns; ++s)
            if (S->sparent[s] >= 0) w[S->sparent[s]] += w[s];  // children precede parents
        double total = 0;
        std::vector<int> roots;
        for (int s = 0; s < ns; ++s)
            if (S->sparent[s] < 0) {
                roots.push_back(s);
                total += w[s];
            }
        std::vector<char> top(ns, 0);
        auto heavier = [&](int a, int b) { return w[a] != w[b] ? w[a] < w[b] : a < b; };  // max-heap on w
        std::make_heap(roots.begin(), roots.end(), heavier);
        // peel the heaviest subtree root into the replicated top until no subtree exceeds its fair share
        while (!roots.empty()) {
            const int r = roots.front();
            const bool too_heavy = w[r] > total / (4.0 * Q);
            const bool too_few = (int)roots.size() < 2 * Q;
            if (!(too_heavy || too_few) || (int)roots.size() >= 64 * Q) break;
            if (S->childptr[r + 1] == S->childptr[r]) break;  // a leaf cannot be split
            std::pop_heap(roots.begin(), roots.end(), heavier);
            roots.pop_back();
            top[r] = 1;
            for (int q = S->childptr[r]; q < S->childptr[r + 1]; ++q) {
                roots.push_back(S->child[q]);
                std::push_heap(roots.begin(), roots.end(), heavier);
            }
        }
        std::sort(roots.begin(), roots.end(), [&](int a, int b) { return w[a] != w[b] ? w[a] > w[b] : a < b; });
        std::vector<double> load(Q, 0.0);
        std::vector<int> assigned(ns, -2);
        for (int r : roots) {
            int best = 0;
            for (int q = 1; q < Q; ++q)
                if (load[q] < load[best]) best = q;
            assigned[r] = best;
            load[best] += w[r];
        }
        for (int s = ns - 1; s >= 0; --s) {
            if (top[s]) S->owner[s] = -1;
            else if (assigned[s] >= 0) S->owner[s] = assigned[s];
            else S->owner[s] = S->owner[S->sparent[s]];
        }
    }
    // ---- update-matrix pool: slot of s lives from level(s) until its parent's level has run ----------
    S->ldu.assign(ns, 0);
    S->uoff.assign(ns, -1);
    auto is_xroot = [&](int s) { return S->owner[s] >= 0 && S->sparent[s] >= 0 && S->owner[S->sparent[s]] < 0; };
    {
        struct Block {
            long long off, size;
        };
        std::vector<Block> freel;  // sorted by offset, coalesced
        long long top = 0;
        auto release = [&](long long off, long long size) {
            auto it = std::lower_bound(freel.begin(), freel.end(), off,
                                       [](const Block& b, long long o) { return b.off < o; });
            it = freel.insert(it, Block{off, size});
            if (it + 1 != freel.end() && it->off + it->size == (it + 1)->off) {
                it->size += (it + 1)->size;
                freel.erase(it + 1);
            }
            if (it != freel.begin() && (it - 1)->off + (it - 1)->size == it->off) {
                (it - 1)->size += it->size;
                it = freel.erase(it) - 1;
            }
            if (it->off + it->size == top) {  // give the tail back
                top = it->off;
                freel.erase(it);
            }
        };
        std::vector<std::vector<int>> free_after(S->nlevels);
        for (int l = 0; l < S->nlevels; ++l) {
            for (int s = S->lvlptr[l]; s < S->lvlptr[l + 1]; ++s) {
                const int nu = S->nr[s] - (first[s + 1] - first[s]);
                if (nu == 0) continue;
                S->ldu[s] = (nu + 15) & ~15;
                if (is_xroot(s)) continue;  // placed in the exchange region below
                const long long need = (long long)S->ldu[s] * nu;
                long long got = -1;
                for (size_t b = 0; b < freel.size(); ++b) {
                    if (freel[b].size >= need) {
                        got = freel[b].off;
                        freel[b].off += need;
                        freel[b].size -= need;
                        if (freel[b].size == 0) freel.erase(freel.begin() + b);
                        break;
                    }
                }
                if (got < 0) {
                    got = top;
                    top += need;
                }
                S->uoff[s] = got;
                S->usize = std::max(S->usize, got + need);
                free_after[S->level[S->sparent[s]]].push_back(s);
            }
            for (int s : free_after[l])
                release(S->uoff[s], (long long)S->ldu[s] * (S->nr[s] - (first[s + 1] - first[s])));
        }
    }

    // exchange regions (multi-GPU) and the update-vector layout of the solves
    {
        const int Q = opt.nranks;
        S->vptr.assign(ns + 1, 0);
        long long vtop = 0;
        if (Q > 1) {
            S->xu_off.assign(Q + 1, 0);
            S->xv_off.assign(Q + 1, 0);
            long long utop = S->usize;
            for (int q = 0; q < Q; ++q) {
                S->xu_off[q] = utop;
                S->xv_off[q] = vtop;
                for (int s = 0; s < ns; ++s) {
                    if (S->owner[s] != q || !is_xroot(s)) continue;
                    const int nu = S->nr[s] - (first[s + 1] - first[s]);
                    S->uoff[s] = utop;
                    utop += (long long)S->ldu[s] * nu;
                    S->vptr[s] = vtop;
                    vtop += (nu + 1) & ~1;
                }
            }
            S->xu_off[Q] = utop;
            S->xv_off[Q] = vtop;
            S->usize = utop;
        }
        for (int s = 0; s < ns; ++s) {
            if (Q > 1 && is_xroot(s)) continue;
            S->vptr[s] = vtop;
            vtop += (S->nr[s] - (first[s + 1] - first[s]) + 1) & ~1;
        }
        S->vptr[ns] = vtop;
        S->vsize = vtop;
    }

    // ---- solve segments ---------------------------------------------------------------------------------------
    S->segptr.assign(ns + 1, 0);
    std::vector<int> seg_src;
    for (int s = 0; s < ns; ++s) {
        const int nc = first[s + 1] - first[s], nu = S->nr[s] - nc;
        const int* R = S->rows.data() + S->rowptr[s] + nc;
        int j = 0;
        while (j < nu) {
            const int t = col2sn[R[j]];
            int j1 = j;
            while (j1 < nu && col2sn[R[j1]] == t) ++j1;
            S->seg_tid.push_back(t);
            S->seg_j0.push_back(j);
            S->seg_j1.push_back(j1);
            seg_src.push_back(s);
            j = j1;
        }
        S->segptr[s + 1] = (int)S->seg_tid.size();
    }
    {
        const int nseg = (int)S->seg_tid.size();
        S->inptr.assign(ns + 1, 0);
        S->in_s.resize(nseg);
        S->in_j0.resize(nseg);
        S->in_j1.resize(nseg);
        for (int q = 0; q < nseg; ++q) S->inptr[S->seg_tid[q] + 1]++;
        for (int t = 0; t < ns; ++t) S->inptr[t + 1] += S->inptr[t];
        std::vector<int> next(S->inptr.begin(), S->inptr.end() - 1);
        for (int q = 0; q < nseg; ++q) {
            const int p = next[S->seg_tid[q]]++;
            S->in_s[p] = seg_src[q];
            S->in_j0[p] = S->seg_j0[q];
            S->in_j1[p] = S->seg_j1[q];
        }
    }

    lap("maps, pool, segments");
    // ---- assembly map -----------------------------------------------------------------------------------------
    {
        // columns are independent: count the entries of every column, prefix-sum, fill in parallel
        std::vector<long long> cstart(m + 1, 0);
        // the three output arrays (anz entries each, 320 MB at config 4) are value-initialised by resize():
        // that sequential fill runs on its own threads beside the counting pass
        std::thread fill_i([&]() { S->ei.resize(S->anz); });
        std::thread fill_j([&]() { S->ej.resize(S->anz); });
        std::thread fill_d([&]() { S->edest.resize(S->anz); });
        parallel_chunks(m, host_threads(), [&](int, int b, int e) {
            for (int j = b; j < e; ++j) {
                const int oj = perm[j];
                int cnt = 1;
                for (int q = ap[oj]; q < ap[oj + 1]; ++q) cnt += iperm[ai[q]] > j;
                cstart[j + 1] = cnt;
            }
        });
        for (int j = 0; j < m; ++j) cstart[j + 1] += cstart[j];
        fill_i.join();
        fill_j.join();
        fill_d.join();
        if (cstart[m] != S->anz)
            return set_err(err, errlen, "symbolic analysis: assembled %d entries, expected %d", (int)cstart[m], (int)S->anz);
        std::atomic<int> bad{-1};
        parallel_chunks(m, host_threads(), [&](int, int b, int e) {
            // the entries of a column leave in row order without sorting: every neighbour is flagged at its
            // position in the supernode's (sorted) row list, then the flags are read in order
            std::vector<int> pos(m, -1), owner_sn(m, -1);
            std::vector<unsigned char> flag;
            int cur = -1;
            for (int j = b; j < e; ++j) {
                const int s = col2sn[j];
                const int* R = S->rows.data() + S->rowptr[s];
                const int nrs = S->nr[s], ncs = first[s + 1] - first[s];
                if (s != cur) {
                    cur = s;
                    for (int p = 0; p < nrs; ++p) {
                        pos[R[p]] = p;
                        owner_sn[R[p]] = s;
                    }
                    flag.assign(nrs, 0);
                }
                const long long base = S->off[s] + (long long)(j - first[s]) * S->ld[s];
                const int oj = perm[j];
                const int pj = j - first[s];
                flag[pj] = 1;
                int lo = pj, hi = pj;
                bool ok = true;
                for (int q = ap[oj]; q < ap[oj + 1]; ++q) {
                    const int i = iperm[ai[q]];
                    if (i <= j) continue;
                    if (owner_sn[i] != s) {
                        ok = false;
                        break;
                    }
                    flag[pos[i]] = 1;
                    hi = std::max(hi, pos[i]);
                }
                if (!ok) {
                    bad = j;
                    for (int p = lo; p < nrs; ++p) flag[p] = 0;
                    continue;
                }
                long long w = cstart[j];
                for (int p = lo; p <= hi; ++p) {
                    if (!flag[p]) continue;
                    flag[p] = 0;
                    S->ei[w] = perm[R[p]];
                    S->ej[w] = oj;
                    S->edest[w] = base + (p < ncs ? p : p + S->nb0[s] - ncs);
                    ++w;
                }
            }
        });
        if (bad >= 0) return set_err(err, errlen, "symbolic analysis: an entry of column %d lies outside its supernode", bad, 0);
    }
    lap("assembly map");
    if ((long long)S->ei.size() != S->anz)
        return set_err(err, errlen, "symbolic analysis: assembled %d entries, expected %d", (int)S->ei.size(), (int)S->anz);

    return 0;
}

}  // namespace nes

// ---- C ABI: the analysis on its own (include/nes.h) ----------------------------------------------------
#include <cstring>

#include "../../include/nes.h"

extern "C" {

void* nes_symbolic_create(int nrow, int ncol, const int* colptr, const int* rowidx, int nranks, int nd_leaf,
                          char* err, size_t errlen) {
    if (nrow < 0 || ncol < 0 || !colptr || (!rowidx && colptr[ncol] > 0)) return nullptr;
    nes::Symbolic* S = new nes::Symbolic();
    nes::SymbolicOptions opt;
    opt.nranks = nranks < 1 ? 1 : nranks;
    opt.nd_leaf = nd_leaf;
    if (nes::symbolic_analyze(nrow, ncol, colptr, rowidx, opt, S, err, errlen) != 0) {
        delete S;
        return nullptr;
    }
    return S;
}

long long nes_symbolic_ints(const void* sym, const char* name, const int** data) {
    const nes::Symbolic* S = static_cast<const nes::Symbolic*>(sym);
    if (!S || !name) return -1;
    const struct {
        const char* n;
        const std::vector<int>* v;
    } tab[] = {{"perm", &S->perm},       {"first", &S->first},     {"nr", &S->nr},           {"ld", &S->ld},
               {"rows", &S->rows},       {"rowptr", &S->rowptr},   {"sparent", &S->sparent}, {"level", &S->level},
               {"lvlptr", &S->lvlptr},   {"childptr", &S->childptr}, {"child", &S->child},   {"relptr", &S->relptr},
               {"nb0", &S->nb0},         {"tbptr", &S->tbptr},     {"tb", &S->tb},
               {"rel", &S->rel},         {"cut", &S->cut},         {"ldu", &S->ldu},         {"owner", &S->owner},
               {"ei", &S->ei},           {"ej", &S->ej},           {"segptr", &S->segptr},   {"seg_tid", &S->seg_tid},
               {"inptr", &S->inptr},     {"in_s", &S->in_s}};
    for (const auto& t : tab)
        if (!strcmp(t.n, name)) {
            if (data) *data = t.v->data();
            return (long long)t.v->size();
        }
    return -1;
}

long long nes_symbolic_longs(const void* sym, const char* name, const long long** data) {
    const nes::Symbolic* S = static_cast<const nes::Symbolic*>(sym);
    if (!S || !name) return -1;
    const std::vector<long long>* v = !strcmp(name, "off") ? &S->off
                                      : !strcmp(name, "uoff") ? &S->uoff
                                      : !strcmp(name, "edest") ? &S->edest
                                      : !strcmp(name, "vptr") ? &S->vptr
                                      : !strcmp(name, "xu_off") ? &S->xu_off
                                      : !strcmp(name, "xv_off") ? &S->xv_off
                                                               : nullptr;
    if (!v) return -1;
    if (data) *data = v->data();
    return (long long)v->size();
}

double nes_symbolic_scalar(const void* sym, const char* name) {
    const nes::Symbolic* S = static_cast<const nes::Symbolic*>(sym);
    if (!S || !name) return -1.0;
    if (!strcmp(name, "anz")) return (double)S->anz;
    if (!strcmp(name, "aatfl")) return S->aatfl;
    if (!strcmp(name, "lnz")) return S->lnz;
    if (!strcmp(name, "fl")) return S->fl;
    if (!strcmp(name, "lsize")) return (double)S->lsize;
    if (!strcmp(name, "usize")) return (double)S->usize;
    if (!strcmp(name, "nsuper")) return (double)S->nsuper;
    if (!strcmp(name, "nlevels")) return (double)S->nlevels;
    return -1.0;
}

void nes_symbolic_free(void* sym) { delete static_cast<nes::Symbolic*>(sym); }

}  // extern "C"
