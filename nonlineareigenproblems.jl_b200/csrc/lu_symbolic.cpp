// Symbolic analysis for the device multifrontal LU: fill-reducing ordering (approximate minimum degree on
// the pattern of A + A^T), elimination tree, postorder, column counts, relaxed supernodes (= fronts),
// front row structures, assembly / extend-add index maps and the level schedule.
//
// This is the integer half of what `factorize(compute_Mder(nep, lambda))` does inside UMFPACK for the
// reference (src/LinSolvers.jl:114-121, src/LinSolverCreators.jl:81); it depends only on the union sparsity
// pattern of the SPMF, so it runs once per operator and is shared by every shift (all N quadrature points of
// contour_beyn, every entry of a factorisation cache).
#include "lu_symbolic.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "common.h"

namespace nepb {

// ------------------------------------------------------------------------------------------------
// approximate minimum degree (quotient graph, approximate external degrees, mass elimination,
// supervariable detection, aggressive element absorption) after Amestoy, Davis & Duff (1996)
// ------------------------------------------------------------------------------------------------
void amd_order(int n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj, std::vector<int32_t>& out) {
    enum : uint8_t { VAR = 0, ELEM = 1, DEAD = 2 };
    std::vector<std::vector<int32_t>> adjV(n), adjE(n), elemVars(n);
    std::vector<int32_t> nv(n, 1), degree(n), elemSize(n, 0), w(n, 0);
    std::vector<uint8_t> status(n, VAR);
    std::vector<int32_t> mark(n, 0), wmark(n, 0), cmark(n, 0);
    std::vector<int32_t> child_head(n, -1), child_next(n, -1), child_tail(n, -1);
    std::vector<int32_t> head(n + 1, -1), nxt(n, -1), prv(n, -1);
    int stamp = 0, wstamp = 0, cstamp = 0;

    auto list_insert = [&](int i) {
        int d = degree[i];
        nxt[i] = head[d];
        prv[i] = -1;
        if (head[d] >= 0) prv[head[d]] = i;
        head[d] = i;
    };
    auto list_remove = [&](int i) {
        if (prv[i] >= 0) nxt[prv[i]] = nxt[i];
        else head[degree[i]] = nxt[i];
        if (nxt[i] >= 0) prv[nxt[i]] = prv[i];
        nxt[i] = prv[i] = -1;
    };
    auto attach = [&](int parent, int c) {  // c is ordered right after parent (and parent's earlier children)
        if (child_head[parent] < 0) child_head[parent] = c;
        else child_next[child_tail[parent]] = c;
        child_tail[parent] = c;
    };

    for (int i = 0; i < n; ++i) {
        adjV[i].assign(adj.begin() + xadj[i], adj.begin() + xadj[i + 1]);
        degree[i] = (int)adjV[i].size();
        list_insert(i);
    }
    std::vector<int32_t> pivots;
    pivots.reserve(n);
    std::vector<int32_t> Lp, newE, survivors;
    std::vector<std::pair<uint32_t, int32_t>> hashes;
    int nel = 0, mindeg = 0;
    while (nel < n) {
        while (mindeg <= n && head[mindeg] < 0) ++mindeg;
        const int p = head[mindeg];
        list_remove(p);
        pivots.push_back(p);
        // ---- form L_p -------------------------------------------------------------------------
        ++stamp;
        mark[p] = stamp;
        Lp.clear();
        for (int v : adjV[p])
            if (status[v] == VAR && mark[v] != stamp) {
                mark[v] = stamp;
                Lp.push_back(v);
            }
        for (int e : adjE[p]) {
            if (status[e] != ELEM) continue;
            for (int v : elemVars[e])
                if (status[v] == VAR && mark[v] != stamp) {
                    mark[v] = stamp;
                    Lp.push_back(v);
                }
            status[e] = DEAD;  // absorbed into p
            std::vector<int32_t>().swap(elemVars[e]);
        }
        std::vector<int32_t>().swap(adjV[p]);
        std::vector<int32_t>().swap(adjE[p]);
        status[p] = ELEM;
        nel += nv[p];
        int degLp = 0;
        for (int v : Lp) degLp += nv[v];
        // ---- w[e] = |L_e \ L_p| for every element adjacent to a variable of L_p ------------------
        ++wstamp;
        for (int i : Lp)
            for (int e : adjE[i]) {
                if (status[e] != ELEM) continue;
                if (wmark[e] != wstamp) {
                    wmark[e] = wstamp;
                    w[e] = elemSize[e];
                }
                w[e] -= nv[i];
            }
        // ---- update the variables of L_p ----------------------------------------------------------
        survivors.clear();
        hashes.clear();
        for (int i : Lp) {
            list_remove(i);
            newE.clear();
            int64_t dege = 0;
            uint32_t h = 0;
            for (int e : adjE[i]) {
                if (status[e] != ELEM) continue;
                if (w[e] <= 0) {  // aggressive absorption: L_e is a subset of L_p
                    status[e] = DEAD;
                    std::vector<int32_t>().swap(elemVars[e]);
                    continue;
                }
                newE.push_back(e);
                dege += w[e];
                h += (uint32_t)e;
            }
            auto& av = adjV[i];
            size_t keep = 0;
            int64_t degv = 0;
            for (int v : av)
                if (status[v] == VAR && mark[v] != stamp) {
                    av[keep++] = v;
                    degv += nv[v];
                    h += (uint32_t)v;
                }
            av.resize(keep);
            if (newE.empty() && av.empty()) {  // mass elimination: indistinguishable from p
                status[i] = DEAD;
                nel += nv[i];
                attach(p, i);
                std::vector<int32_t>().swap(adjE[i]);
                continue;
            }
            newE.push_back(p);
            h += (uint32_t)p;
            adjE[i].assign(newE.begin(), newE.end());
            int64_t d = degv + dege + (degLp - nv[i]);
            d = std::min<int64_t>(d, (int64_t)degree[i] + degLp - nv[i]);
            degree[i] = (int)d;  // clamped against n - nel below, after mass elimination finished
            survivors.push_back(i);
            hashes.emplace_back(h, i);
        }
        // ---- supervariable detection ---------------------------------------------------------------
        std::sort(hashes.begin(), hashes.end());
        for (size_t a = 0; a < hashes.size();) {
            size_t b = a;
            while (b < hashes.size() && hashes[b].first == hashes[a].first) ++b;
            for (size_t x = a; x < b; ++x) {
                const int i = hashes[x].second;
                if (status[i] != VAR) continue;
                bool marked = false;
                for (size_t y = x + 1; y < b; ++y) {
                    const int j = hashes[y].second;
                    if (status[j] != VAR) continue;
                    if (adjV[i].size() != adjV[j].size() || adjE[i].size() != adjE[j].size()) continue;
                    if (!marked) {
                        ++cstamp;
                        for (int v : adjV[i]) cmark[v] = cstamp;
                        for (int e : adjE[i]) cmark[e] = cstamp;
                        marked = true;
                    }
                    bool same = true;
                    for (int v : adjV[j])
                        if (cmark[v] != cstamp) { same = false; break; }
                    if (same)
                        for (int e : adjE[j])
                            if (cmark[e] != cstamp) { same = false; break; }
                    if (!same) continue;
                    nv[i] += nv[j];
                    degree[i] -= nv[j];
                    nv[j] = 0;
                    status[j] = DEAD;
                    attach(i, j);
                    std::vector<int32_t>().swap(adjV[j]);
                    std::vector<int32_t>().swap(adjE[j]);
                }
            }
            a = b;
        }
        // ---- finalise element p ---------------------------------------------------------------------
        auto& ev = elemVars[p];
        ev.clear();
        int size = 0;
        for (int i : survivors)
            if (status[i] == VAR) {
                ev.push_back(i);
                size += nv[i];
            }
        elemSize[p] = size;
        if (ev.empty()) status[p] = DEAD;
        for (int i : ev) {
            int d = std::min(degree[i], n - nel - nv[i]);
            degree[i] = std::max(d, 0);
            list_insert(i);
            if (degree[i] < mindeg) mindeg = degree[i];
        }
    }
    // ---- emit: every pivot followed by the variables eliminated with it (depth first) -----------------
    out.clear();
    out.reserve(n);
    std::vector<int32_t> stack;
    for (int p : pivots) {
        stack.push_back(p);
        while (!stack.empty()) {
            int v = stack.back();
            stack.pop_back();
            out.push_back(v);
            // push children in reverse so that they come out in attach order
            size_t base = stack.size();
            for (int c = child_head[v]; c >= 0; c = child_next[c]) stack.push_back(c);
            std::reverse(stack.begin() + base, stack.end());
        }
    }
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
// strict lower-triangular pattern of P (A + A^T) P^T by rows: for every row i the columns k < i, ascending
static void permuted_lower(int n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj,
                           const std::vector<int32_t>& perm, const std::vector<int32_t>& iperm,
                           std::vector<int64_t>& lptr, std::vector<int32_t>& lcol) {
    lptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        const int o = perm[i];
        int c = 0;
        for (int64_t e = xadj[o]; e < xadj[o + 1]; ++e)
            if (iperm[adj[e]] < i) ++c;
        lptr[i + 1] = lptr[i] + c;
    }
    lcol.resize(lptr[n]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        const int o = perm[i];
        int64_t t = lptr[i];
        for (int64_t e = xadj[o]; e < xadj[o + 1]; ++e) {
            const int k = iperm[adj[e]];
            if (k < i) lcol[t++] = k;
        }
        std::sort(lcol.begin() + lptr[i], lcol.begin() + lptr[i + 1]);
    }
}

static void etree_liu(int n, const std::vector<int64_t>& lptr, const std::vector<int32_t>& lcol, std::vector<int32_t>& parent) {
    parent.assign(n, -1);
    std::vector<int32_t> anc(n, -1);
    for (int i = 0; i < n; ++i)
        for (int64_t e = lptr[i]; e < lptr[i + 1]; ++e) {
            int k = lcol[e];
            while (k != -1 && k < i) {
                const int next = anc[k];
                anc[k] = i;
                if (next == -1) parent[k] = i;
                k = next;
            }
        }
}

static void postorder(int n, const std::vector<int32_t>& parent, std::vector<int32_t>& post) {
    std::vector<int32_t> head(n, -1), next(n, -1);
    for (int j = n - 1; j >= 0; --j)
        if (parent[j] >= 0) {
            next[j] = head[parent[j]];
            head[parent[j]] = j;
        }
    post.clear();
    post.reserve(n);
    std::vector<int32_t> stack;
    for (int r = 0; r < n; ++r) {
        if (parent[r] >= 0) continue;
        stack.push_back(r);
        while (!stack.empty()) {
            const int v = stack.back();
            const int c = head[v];
            if (c < 0) {
                post.push_back(v);
                stack.pop_back();
            } else {
                head[v] = next[c];
                stack.push_back(c);
            }
        }
    }
}

namespace {
struct PhaseTimer {  // NEPB_LU_TIMING=1: wall time of every phase of the analysis on stderr
    bool on = getenv("NEPB_LU_TIMING") && atoi(getenv("NEPB_LU_TIMING")) != 0;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[lu_symbolic] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

int lu_symbolic_analyse(int n, const int32_t* rowptr, const int32_t* colind, const int32_t* user_perm,
                        const LuOptions& opt, LuSymbolic& S, const int32_t* rowmap) {
    PhaseTimer timer;
    S = LuSymbolic();
    S.n = n;
    S.nnz = rowptr[n];
    // rowmap (optional, from max_product_matching): row i of the operator is row rowmap[i] of the matrix that is factorised, so
    // that its diagonal is the matching; everything below works on that row-permuted pattern
    if (rowmap) S.rowmap.assign(rowmap, rowmap + n);
    // ---- adjacency of A + A^T without the diagonal ---------------------------------------------------
    std::vector<int64_t> xadj(n + 1, 0);
    std::vector<int32_t> adj;
    {
        std::vector<int32_t> cnt(n, 0);
        for (int i0 = 0; i0 < n; ++i0)
            for (int e = rowptr[i0]; e < rowptr[i0 + 1]; ++e) {
                const int i = rowmap ? rowmap[i0] : i0;
                const int j = colind[e];
                if (j != i) {
                    cnt[i]++;
                    cnt[j]++;
                }
            }
        for (int i = 0; i < n; ++i) xadj[i + 1] = xadj[i] + cnt[i];
        adj.resize(xadj[n]);
        std::vector<int64_t> fill(xadj.begin(), xadj.end() - 1);
        for (int i0 = 0; i0 < n; ++i0)
            for (int e = rowptr[i0]; e < rowptr[i0 + 1]; ++e) {
                const int i = rowmap ? rowmap[i0] : i0;
                const int j = colind[e];
                if (j != i) {
                    adj[fill[i]++] = j;
                    adj[fill[j]++] = i;
                }
            }
        // sort + unique per vertex, then compact
        std::vector<int64_t> nx(n + 1, 0);
#pragma omp parallel for schedule(dynamic, 1024)
        for (int i = 0; i < n; ++i) {
            auto b = adj.begin() + xadj[i], e = adj.begin() + xadj[i + 1];
            std::sort(b, e);
            nx[i + 1] = std::unique(b, e) - b;
        }
        int64_t t = 0;
        for (int i = 0; i < n; ++i) {
            const int64_t len = nx[i + 1], src = xadj[i];
            nx[i] = t;
            for (int64_t x = 0; x < len; ++x) adj[t + x] = adj[src + x];
            t += len;
        }
        nx[n] = t;
        adj.resize(t);
        xadj.swap(nx);
    }
    timer.lap("adjacency of A+A^T");
    // ---- fill-reducing ordering ----------------------------------------------------------------------
    std::vector<int32_t> perm0;
    if (user_perm) {
        perm0.assign(user_perm, user_perm + n);
        std::vector<char> seen(n, 0);
        for (int i = 0; i < n; ++i) {
            NEPB_CHECK_ARG(perm0[i] >= 0 && perm0[i] < n && !seen[perm0[i]], "user permutation is not a permutation of 0..n-1");
            seen[perm0[i]] = 1;
        }
    } else if (opt.ordering == 1) {
        perm0.resize(n);
        std::iota(perm0.begin(), perm0.end(), 0);
    } else {
        amd_order(n, xadj, adj, perm0);
    }
    timer.lap("ordering");
    NEPB_CHECK_ARG((int)perm0.size() == n, "ordering produced %d of %d vertices", (int)perm0.size(), n);
    std::vector<int32_t> iperm0(n);
    for (int i = 0; i < n; ++i) iperm0[perm0[i]] = i;
    // ---- etree, postorder, compose -------------------------------------------------------------------
    std::vector<int64_t> lptr;
    std::vector<int32_t> lcol, parent0, post;
    permuted_lower(n, xadj, adj, perm0, iperm0, lptr, lcol);
    etree_liu(n, lptr, lcol, parent0);
    postorder(n, parent0, post);
    S.perm.resize(n);
    S.iperm.resize(n);
    for (int i = 0; i < n; ++i) S.perm[i] = perm0[post[i]];
    for (int i = 0; i < n; ++i) S.iperm[S.perm[i]] = i;
    permuted_lower(n, xadj, adj, S.perm, S.iperm, lptr, lcol);
    etree_liu(n, lptr, lcol, S.parent);
    const auto& parent = S.parent;
    for (int j = 0; j < n; ++j) NEPB_CHECK_ARG(parent[j] == -1 || parent[j] > j, "internal: etree is not postordered");
    timer.lap("etree + postorder");
    // ---- column counts ------------------------------------------------------------------------------------------
    // column-wise lower pattern of the permuted matrix: rows i > k of column k, ascending (lptr / lcol is its row-wise form)
    std::vector<int64_t> cptr(n + 1, 0);
    std::vector<int32_t> crow(lptr[n]);
    {
        for (int64_t e = 0; e < lptr[n]; ++e) cptr[lcol[e] + 1]++;
        for (int k = 0; k < n; ++k) cptr[k + 1] += cptr[k];
        std::vector<int64_t> nxt(cptr.begin(), cptr.end() - 1);
        for (int i = 0; i < n; ++i)
            for (int64_t e = lptr[i]; e < lptr[i + 1]; ++e) crow[nxt[lcol[e]]++] = i;
    }
    const bool legacy = getenv("NEPB_LU_LEGACY") && atoi(getenv("NEPB_LU_LEGACY")) != 0;   // the O(nnz(L)) traversals
    const bool cross_check = getenv("NEPB_LU_CHECK") && atoi(getenv("NEPB_LU_CHECK")) != 0;  // run both, insist on equality
    auto counts_by_traversal = [&](std::vector<int32_t>& colcount) {
        colcount.assign(n, 1);
        std::vector<int32_t> vis(n, -1);
        for (int i = 0; i < n; ++i) {
            vis[i] = i;
            for (int64_t e = lptr[i]; e < lptr[i + 1]; ++e)
                for (int k = lcol[e]; vis[k] != i; k = parent[k]) {
                    vis[k] = i;
                    colcount[k]++;
                }
        }
    };
    // Gilbert / Ng / Peyton: count the leaves of every row subtree through the skeleton of the matrix; an entry (i, j), i > j,
    // is in the skeleton when j is a leaf of the row subtree of i, recognised from first descendants in the (identity)
    // postorder; overlaps of consecutive leaves are charged to their least common ancestor found by path compression.
    // O(nnz(A) alpha(n)) instead of O(nnz(L)).
    auto counts_by_skeleton = [&](std::vector<int32_t>& colcount) {
        std::vector<int32_t> first(n, -1), maxfirst(n, -1), prevleaf(n, -1), ancestor(n);
        std::vector<int32_t>& delta = colcount;
        delta.assign(n, 0);
        for (int k = 0; k < n; ++k) {  // columns are postordered: post[k] = k
            int j = k;
            delta[j] = (first[j] == -1) ? 1 : 0;  // j is a leaf of the elimination tree
            for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
        }
        for (int i = 0; i < n; ++i) ancestor[i] = i;
        for (int j = 0; j < n; ++j) {
            if (parent[j] != -1) delta[parent[j]]--;  // j is not a root
            for (int64_t e = cptr[j]; e < cptr[j + 1]; ++e) {
                const int i = crow[e];  // i > j
                if (first[j] <= maxfirst[i]) continue;  // j is not a leaf of the row subtree of i
                maxfirst[i] = first[j];
                const int jprev = prevleaf[i];
                prevleaf[i] = j;
                delta[j]++;  // (i, j) is in the skeleton
                if (jprev != -1) {
                    int q = jprev;
                    while (q != ancestor[q]) q = ancestor[q];
                    for (int t = jprev; t != q;) {
                        const int tp = ancestor[t];
                        ancestor[t] = q;
                        t = tp;
                    }
                    delta[q]--;  // the overlap of the two leaf-to-i paths starts at their least common ancestor
                }
            }
            if (parent[j] != -1) ancestor[j] = parent[j];
        }
        for (int j = 0; j < n; ++j)
            if (parent[j] != -1) colcount[parent[j]] += colcount[j];  // parent[j] > j
    };
    if (legacy) counts_by_traversal(S.colcount);
    else counts_by_skeleton(S.colcount);
    if (cross_check) {
        std::vector<int32_t> cc2;
        if (legacy) counts_by_skeleton(cc2);
        else counts_by_traversal(cc2);
        NEPB_CHECK_ARG(cc2 == S.colcount, "internal: the two column-count algorithms disagree");
    }
    const auto& cc = S.colcount;
    timer.lap("column counts");
    // ---- supernode partition ----------------------------------------------------------------------------
    std::vector<int32_t> subtree(n, 1);
    for (int j = 0; j < n; ++j)
        if (parent[j] >= 0) subtree[parent[j]] += subtree[j];
    const int max_np = std::max(1, opt.max_np);
    const int relax = std::min(opt.relax_leaf, max_np);
    std::vector<char> starts(n, 0);  // 1 = column begins a supernode
    {
        // (a) maximal small subtrees -> one supernode each
        std::vector<int32_t> small_root(n, -1);  // for columns inside a small subtree: its root
        for (int j = n - 1; j >= 0; --j) {
            if (small_root[j] >= 0) continue;
            if (subtree[j] <= relax) {
                for (int c = j - subtree[j] + 1; c <= j; ++c) small_root[c] = j;  // postorder: subtree is contiguous
            }
        }
        int first = 0;       // first column of the current supernode
        double dense = 0, truennz = 0;
        for (int j = 0; j < n; ++j) {
            bool merge = false;
            if (j > 0 && j - first < max_np) {
                if (small_root[j] >= 0 && small_root[j] == small_root[j - 1]) {
                    merge = true;
                } else if (parent[j - 1] == j && small_root[j] < 0) {
                    // chain merge with a bound on the explicit zeros (thresholds as in CHOLMOD's relaxed supernodes)
                    const int wnew = j - first + 1;
                    const double nfnew = wnew + cc[j] - 1;
                    double d = 0;
                    for (int t = 0; t < wnew; ++t) d += nfnew - t;
                    const double tn = truennz + cc[j];
                    const double z = (d - tn) / d;
                    if (cc[j - 1] == cc[j] + 1 && dense == truennz) merge = true;  // fundamental: no zeros at all
                    else if (wnew <= 4) merge = true;
                    else if (wnew <= 16) merge = z < 0.8;
                    else if (wnew <= 48) merge = z < 0.1;
                    else merge = z < 0.05;
                }
            }
            if (!merge) {
                starts[j] = 1;
                first = j;
                truennz = 0;
            }
            truennz += cc[j];
            const int wcur = j - first + 1;
            const double nf = wcur + cc[j] - 1;
            dense = 0;
            for (int t = 0; t < wcur; ++t) dense += nf - t;
        }
    }
    S.col_sn.resize(n);
    S.sn_ptr.clear();
    for (int j = 0; j < n; ++j) {
        if (starts[j]) S.sn_ptr.push_back(j);
        S.col_sn[j] = (int)S.sn_ptr.size() - 1;
    }
    S.nsuper = (int)S.sn_ptr.size();
    S.sn_ptr.push_back(n);
    const int ns = S.nsuper;
    S.sn_parent.assign(ns, -1);
    for (int s = 0; s < ns; ++s) {
        // inside a relaxed subtree supernode the last column is the subtree root, otherwise the chain end
        const int last = S.sn_ptr[s + 1] - 1;
        int pj = parent[last];
        // columns of a small subtree may have parents inside s; the supernode's parent is the first ancestor outside
        while (pj >= 0 && S.col_sn[pj] == s) pj = parent[pj];
        S.sn_parent[s] = pj >= 0 ? S.col_sn[pj] : -1;
    }
    timer.lap("supernodes");
    // ---- front row structures ----------------------------------------------------------------------------------
    // rows(s) = pivot columns of s, then (sorted) every row i beyond them with L(i, k) != 0 for some column k of s.
    // Default: supernodal symbolic factorisation -- the update rows of s are the entries of A below its columns merged with
    // the update rows of its children (children precede parents: the columns are postordered), O(sum of front orders) work.
    // NEPB_LU_LEGACY=1 selects the original row-subtree traversals (O(nnz(L)): column counts + two passes for the structures
    // took 3.1 of 6.0 s at n = 10^6); NEPB_LU_CHECK=1 computes both variants and insists that they are identical.
    auto rows_by_traversal = [&](std::vector<int64_t>& row_ptr, std::vector<int32_t>& rows) {
        std::vector<int64_t> cnt(ns, 0);
        std::vector<int32_t> vis(n, -1), snvis(ns, -1);
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                row_ptr.assign(ns + 1, 0);
                for (int s = 0; s < ns; ++s) row_ptr[s + 1] = row_ptr[s] + (S.sn_ptr[s + 1] - S.sn_ptr[s]) + cnt[s];
                rows.resize(row_ptr[ns]);
                for (int s = 0; s < ns; ++s) {
                    int64_t t = row_ptr[s];
                    for (int c = S.sn_ptr[s]; c < S.sn_ptr[s + 1]; ++c) rows[t++] = c;
                    cnt[s] = t;  // write cursor
                }
                std::fill(vis.begin(), vis.end(), -1);
                std::fill(snvis.begin(), snvis.end(), -1);
            }
            for (int i = 0; i < n; ++i) {
                vis[i] = i;
                const int si = S.col_sn[i];
                for (int64_t e = lptr[i]; e < lptr[i + 1]; ++e)
                    for (int k = lcol[e]; vis[k] != i; k = parent[k]) {
                        vis[k] = i;
                        const int s = S.col_sn[k];
                        if (s != si && snvis[s] != i) {
                            snvis[s] = i;
                            if (pass == 0) cnt[s]++;
                            else rows[cnt[s]++] = i;
                        }
                    }
            }
        }
    };
    auto rows_by_children = [&](std::vector<int64_t>& row_ptr, std::vector<int32_t>& rows) {
        std::vector<int32_t> cp(ns + 1, 0), cl(ns);
        for (int s = 0; s < ns; ++s)
            if (S.sn_parent[s] >= 0) cp[S.sn_parent[s] + 1]++;
        for (int s = 0; s < ns; ++s) cp[s + 1] += cp[s];
        {
            std::vector<int32_t> nxt(cp.begin(), cp.end() - 1);
            for (int s = 0; s < ns; ++s)
                if (S.sn_parent[s] >= 0) cl[nxt[S.sn_parent[s]]++] = s;
        }
        row_ptr.assign(ns + 1, 0);
        rows.clear();
        rows.reserve((size_t)n * 4);
        std::vector<int32_t> mark(n, -1), cur;
        for (int s = 0; s < ns; ++s) {
            const int first = S.sn_ptr[s], last = S.sn_ptr[s + 1] - 1;
            cur.clear();
            for (int k = first; k <= last; ++k)
                for (int64_t e = cptr[k]; e < cptr[k + 1]; ++e) {
                    const int i = crow[e];
                    if (i > last && mark[i] != s) {
                        mark[i] = s;
                        cur.push_back(i);
                    }
                }
            for (int ci = cp[s]; ci < cp[s + 1]; ++ci) {
                const int c = cl[ci];
                const int64_t b0 = row_ptr[c] + (S.sn_ptr[c + 1] - S.sn_ptr[c]), b1 = row_ptr[c + 1];
                for (int64_t x = b0; x < b1; ++x) {
                    const int i = rows[x];
                    if (i > last && mark[i] != s) {
                        mark[i] = s;
                        cur.push_back(i);
                    }
                }
            }
            std::sort(cur.begin(), cur.end());
            for (int c = first; c <= last; ++c) rows.push_back(c);
            rows.insert(rows.end(), cur.begin(), cur.end());
            row_ptr[s + 1] = (int64_t)rows.size();
        }
    };
    {
        const bool traversal = legacy, check = cross_check;
        if (traversal) rows_by_traversal(S.row_ptr, S.rows);
        else rows_by_children(S.row_ptr, S.rows);
        if (check) {
            std::vector<int64_t> rp2;
            std::vector<int32_t> rows2;
            if (traversal) rows_by_children(rp2, rows2);
            else rows_by_traversal(rp2, rows2);
            NEPB_CHECK_ARG(rp2 == S.row_ptr && rows2 == S.rows, "internal: the two front-structure algorithms disagree");
        }
    }
    timer.lap("front row structures");
    // ---- offsets, statistics, relative indices ---------------------------------------------------------------
    // A front whose only child hands it a contribution block with exactly its own row structure (the links of a
    // supernode split by max_np, and any other such pair) lives *inside* the child's front: its storage is the child's
    // F22 block, so the child's Schur update lands in place and no extend-add / extra memory is needed.
    S.front_off.assign(ns + 1, 0);
    S.front_ld.assign(ns, 0);
    S.in_place_child.assign(ns, 0);
    S.w_off.assign(ns + 1, 0);
    S.rel_ptr.assign(ns + 1, 0);
    {
        std::vector<int32_t> nchild(ns, 0), only_child(ns, -1);
        for (int s = 0; s < ns; ++s)
            if (S.sn_parent[s] >= 0) {
                nchild[S.sn_parent[s]]++;
                only_child[S.sn_parent[s]] = s;
            }
        int64_t total = 0;
        for (int s = 0; s < ns; ++s) {
            const int64_t nf = S.row_ptr[s + 1] - S.row_ptr[s];
            const int c = only_child[s];
            bool alias = false;
            if (opt.alias_chains && nchild[s] == 1) {
                const int64_t nfc = S.row_ptr[c + 1] - S.row_ptr[c];
                const int64_t npc = S.sn_ptr[c + 1] - S.sn_ptr[c];
                alias = (nfc - npc == nf);
                if (alias) {
                    S.front_ld[s] = S.front_ld[c];
                    S.front_off[s] = S.front_off[c] + npc * ((int64_t)S.front_ld[c] + 1);
                    S.in_place_child[c] = 1;
                }
            }
            if (!alias) {
                S.front_ld[s] = (int32_t)nf;
                S.front_off[s] = total;
                total += nf * nf;
            }
        }
        S.front_off[ns] = total;
    }
    // solve work rows: one nf x k block per front; an in-place parent re-uses its child's update rows
    S.has_in_place_child.assign(ns, 0);
    {
        int64_t wtotal = 0;
        std::vector<int32_t> ipc(ns, -1);
        for (int s = 0; s < ns; ++s)
            if (S.in_place_child[s]) ipc[S.sn_parent[s]] = s;
        for (int s = 0; s < ns; ++s) {
            const int64_t nf = S.row_ptr[s + 1] - S.row_ptr[s];
            if (ipc[s] >= 0) {
                const int c = ipc[s];
                S.has_in_place_child[s] = 1;
                S.w_off[s] = S.w_off[c] + (S.sn_ptr[c + 1] - S.sn_ptr[c]);
            } else {
                S.w_off[s] = wtotal;
                wtotal += nf;
            }
        }
        S.w_off[ns] = wtotal;
    }
    for (int s = 0; s < ns; ++s) {
        const int64_t nf = S.row_ptr[s + 1] - S.row_ptr[s];
        const int64_t np = S.sn_ptr[s + 1] - S.sn_ptr[s];
        S.rel_ptr[s + 1] = S.rel_ptr[s] + (nf - np);
        S.nnz_factor += 2 * np * nf - np * np;
        S.max_nf = std::max<int>(S.max_nf, (int)nf);
        S.max_np = std::max<int>(S.max_np, (int)np);
        for (int64_t t = 0; t < np; ++t) {
            const double r = (double)(nf - t - 1);
            S.flops += r + r * r;
        }
    }
    S.front_total = S.front_off[ns];
    S.w_total = S.w_off[ns];
    S.rel.resize(S.rel_ptr[ns]);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(| : bad)
    for (int s = 0; s < ns; ++s) {
        const int ps = S.sn_parent[s];
        const int64_t np = S.sn_ptr[s + 1] - S.sn_ptr[s];
        const int32_t* r = S.rows.data() + S.row_ptr[s] + np;
        const int64_t ncb = S.rel_ptr[s + 1] - S.rel_ptr[s];
        if (ncb == 0) continue;
        if (ps < 0) { bad |= 1; continue; }
        const int32_t* pr = S.rows.data() + S.row_ptr[ps];
        const int64_t pnf = S.row_ptr[ps + 1] - S.row_ptr[ps];
        int64_t t = 0;
        for (int64_t x = 0; x < ncb; ++x) {
            while (t < pnf && pr[t] < r[x]) ++t;
            if (t >= pnf || pr[t] != r[x]) { bad |= 2; break; }
            S.rel[S.rel_ptr[s] + x] = (int32_t)t;
        }
    }
    NEPB_CHECK_ARG(!bad, "internal: update rows of a front are not contained in its parent (code %d)", bad);
    timer.lap("offsets + relative indices");
    // ---- levels ---------------------------------------------------------------------------------------------
    S.level.assign(ns, 0);
    for (int s = 0; s < ns; ++s)
        if (S.sn_parent[s] >= 0) S.level[S.sn_parent[s]] = std::max(S.level[S.sn_parent[s]], S.level[s] + 1);
    S.nlevels = 0;
    for (int s = 0; s < ns; ++s) S.nlevels = std::max(S.nlevels, S.level[s] + 1);
    S.level_ptr.assign(S.nlevels + 1, 0);
    for (int s = 0; s < ns; ++s) S.level_ptr[S.level[s] + 1]++;
    for (int l = 0; l < S.nlevels; ++l) S.level_ptr[l + 1] += S.level_ptr[l];
    S.level_list.resize(ns);
    {
        std::vector<int32_t> cur(S.level_ptr.begin(), S.level_ptr.end() - 1);
        for (int s = 0; s < ns; ++s) S.level_list[cur[S.level[s]]++] = s;
    }
    // ---- assembly map: CSR nonzero (i, j) -> front of min(pi, pj) -----------------------------------------------
    S.a_pos.resize(S.nnz);
    int bad2 = 0;
#pragma omp parallel for schedule(static) reduction(| : bad2)
    for (int i = 0; i < n; ++i) {
        const int pi = S.iperm[rowmap ? rowmap[i] : i];
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int pj = S.iperm[colind[e]];
            const int s = S.col_sn[std::min(pi, pj)];
            const int32_t* r = S.rows.data() + S.row_ptr[s];
            const int64_t nf = S.row_ptr[s + 1] - S.row_ptr[s];
            const int32_t* ri = std::lower_bound(r, r + nf, pi);
            const int32_t* rj = std::lower_bound(r, r + nf, pj);
            if (ri == r + nf || *ri != pi || rj == r + nf || *rj != pj) { bad2 = 1; continue; }
            S.a_pos[e] = S.front_off[s] + (ri - r) + (rj - r) * (int64_t)S.front_ld[s];  // column-major front
        }
    }
    NEPB_CHECK_ARG(!bad2, "internal: a matrix entry does not fall into its front");
    timer.lap("levels + assembly map");
    return NEPB_OK;
}

}  // namespace nepb
