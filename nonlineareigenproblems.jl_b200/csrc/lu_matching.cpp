// Maximum-product row matching with dual scalings, for static pivoting in the device LU.
//
// The reference gets numerical pivoting from UMFPACK (`factorize(compute_Mder(nep, lambda))`, src/LinSolvers.jl:114-121),
// which may pick a pivot anywhere in a column.  The device multifrontal LU only pivots inside the <= 32 x 32 pivot block
// of a front, which is safe when the diagonal carries weight (gun, the stencil PEPs) and fails when it does not (qdep0 at
// sigma = 0, the configuration of test/infbilanczos.jl: 980 of 1000 diagonal entries vanish).  For those operators the rows
// are permuted beforehand so that the product of the diagonal magnitudes is maximal, and rows and columns are scaled with the
// dual variables of that assignment problem so that every matched entry has modulus 1 and every other entry at most 1
// (the I-matrix scaling of Duff & Koster, "On algorithms for permuting large entries to the diagonal of a sparse matrix",
// SIAM J. Matrix Anal. Appl. 22 (2001); Olschowka & Neumaier 1996) -- the published algorithm, written from the paper.
//
// Sparse shortest-augmenting-path assignment: costs c(i,j) = log(max_j |a_ij|) - log|a_ij| >= 0 on the stored entries,
// row duals u, column duals v with c(i,j) - u_i - v_j >= 0 and equality on matched entries; every unmatched row grows a
// Dijkstra tree of alternating paths over reduced costs until it reaches a free column.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <utility>
#include <vector>

#include "common.h"
#include "lu_symbolic.h"

namespace nepb {

// rowptr / colind: CSR pattern (0-based), absval[e] = |a_e|.  Out: row_to_col[i] = column matched to row i,
// dr[i], dc[j] scalings with dr_i |a_ij| dc_j <= 1 (= 1 on the matching).  Returns the number of matched rows (n = success).
int max_product_matching(int n, const int32_t* rowptr, const int32_t* colind, const double* absval,
                         std::vector<int32_t>& row_to_col, std::vector<double>& dr, std::vector<double>& dc) {
    const double INF = std::numeric_limits<double>::infinity();
    const int64_t nnz = rowptr[n];
    std::vector<double> cost(nnz), lrmax(n, 0.0);
    for (int i = 0; i < n; ++i) {
        double m = 0.0;
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) m = std::max(m, absval[e]);
        lrmax[i] = m > 0.0 ? std::log(m) : 0.0;
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) cost[e] = absval[e] > 0.0 ? lrmax[i] - std::log(absval[e]) : INF;
    }
    std::vector<double> u(n, 0.0), v(n, INF);
    std::vector<int32_t> mrow(n, -1), mcol(n, -1);  // mrow[i] = column of row i, mcol[j] = row of column j
    // initial duals: u_i = 0 (every row has a zero-cost entry), v_j = min_i c(i,j); then a greedy pass over tight entries
    for (int i = 0; i < n; ++i)
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) v[colind[e]] = std::min(v[colind[e]], cost[e]);
    for (int j = 0; j < n; ++j)
        if (v[j] == INF) v[j] = 0.0;  // empty column: the matching will fail below
    for (int i = 0; i < n; ++i)
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int j = colind[e];
            if (mcol[j] < 0 && cost[e] - v[j] <= 0.0) {
                mrow[i] = j;
                mcol[j] = i;
                break;
            }
        }
    std::vector<double> dist(n, INF);
    std::vector<int32_t> pred(n, -1), scanned, touched;
    std::vector<char> done(n, 0);
    using QE = std::pair<double, int32_t>;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> heap;
    int matched = 0;
    for (int i = 0; i < n; ++i) matched += mrow[i] >= 0;
    for (int root = 0; root < n; ++root) {
        if (mrow[root] >= 0) continue;
        while (!heap.empty()) heap.pop();
        scanned.clear();
        touched.clear();
        int i = root, jend = -1;
        double lsp = 0.0;
        for (;;) {
            for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
                const int j = colind[e];
                if (done[j] || cost[e] == INF) continue;
                const double nd = lsp + (cost[e] - u[i] - v[j]);
                if (nd < dist[j]) {
                    if (dist[j] == INF) touched.push_back(j);
                    dist[j] = nd;
                    pred[j] = i;
                    heap.push(QE(nd, j));
                }
            }
            int j = -1;
            while (!heap.empty()) {
                const QE t = heap.top();
                heap.pop();
                if (!done[t.second] && t.first <= dist[t.second]) {
                    j = t.second;
                    break;
                }
            }
            if (j < 0) break;  // no augmenting path: structurally singular
            done[j] = 1;
            scanned.push_back(j);
            lsp = dist[j];
            if (mcol[j] < 0) {
                jend = j;
                break;
            }
            i = mcol[j];
        }
        if (jend >= 0) {
            // duals first (they need the old mates), then flip the path
            u[root] += lsp;
            for (int j : scanned) {
                if (j != jend) u[mcol[j]] += lsp - dist[j];
                v[j] -= lsp - dist[j];
            }
            for (int j = jend;;) {
                const int r = pred[j];
                const int jn = mrow[r];
                mrow[r] = j;
                mcol[j] = r;
                if (r == root) break;
                j = jn;
            }
            ++matched;
        }
        for (int j : touched) {
            dist[j] = INF;
            done[j] = 0;
            pred[j] = -1;
        }
    }
    row_to_col = mrow;
    dr.assign(n, 1.0);
    dc.assign(n, 1.0);
    if (matched == n)
        for (int i = 0; i < n; ++i) {
            dr[i] = std::exp(u[i] - lrmax[i]);
            dc[i] = std::exp(v[i]);
        }
    return matched;
}

}  // namespace nepb
