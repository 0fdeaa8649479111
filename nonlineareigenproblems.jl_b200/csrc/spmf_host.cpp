// Host-side integer work of the SPMF operator: union sparsity pattern of the p CSC matrices,
// CSC -> CSR, int64 -> int32, value interleaving.  This is the B200-side equivalent of
// form_aligned_sparsity_patterns (reference src/NEPTypes.jl:244-274), done once per operator.
// Pure integer/merge work -- results are bit-exact and checked against the oracle in tests.
#include <algorithm>
#include <cstring>
#include <limits>
#include <vector>

#include "common.h"
#include "spmf_host.h"

namespace nepb {

int build_union_csr(int64_t n, int p, const int64_t* const* colptr, const int64_t* const* rowval,
                    const void* const* nzval, int is_complex, int base, UnionCSR& u) {
    const int vs = is_complex ? 2 : 1;
    const int vw = p * vs;
    // ---- validation --------------------------------------------------------------------------
    for (int i = 0; i < p; ++i) {
        NEPB_CHECK_ARG(colptr[i] && (rowval[i] || colptr[i][n] == base) && (nzval[i] || colptr[i][n] == base),
                       "matrix %d: NULL array", i);
        NEPB_CHECK_ARG(colptr[i][0] == base, "matrix %d: colptr[0]=%lld, expected index base %d", i,
                       (long long)colptr[i][0], base);
    }
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t j = 0; j < n; ++j) {
        for (int i = 0; i < p; ++i) {
            int64_t lo = colptr[i][j] - base, hi = colptr[i][j + 1] - base;
            if (hi < lo) { bad |= 1; continue; }
            for (int64_t e = lo; e < hi; ++e) {
                int64_t r = rowval[i][e] - base;
                if (r < 0 || r >= n) bad |= 2;
                if (e > lo && rowval[i][e - 1] >= rowval[i][e]) bad |= 4;
            }
        }
    }
    NEPB_CHECK_ARG(!(bad & 1), "colptr is not non-decreasing");
    NEPB_CHECK_ARG(!(bad & 2), "row index out of range");
    NEPB_CHECK_ARG(!(bad & 4), "row indices must be strictly ascending within a column (SparseMatrixCSC invariant)");

    // ---- pass 1: union size per column --------------------------------------------------------
    std::vector<int64_t> ucolptr(n + 1, 0);
#pragma omp parallel
    {
        std::vector<int64_t> tmp;
#pragma omp for schedule(static)
        for (int64_t j = 0; j < n; ++j) {
            tmp.clear();
            for (int i = 0; i < p; ++i)
                for (int64_t e = colptr[i][j] - base; e < colptr[i][j + 1] - base; ++e) tmp.push_back(rowval[i][e] - base);
            std::sort(tmp.begin(), tmp.end());
            ucolptr[j + 1] = std::unique(tmp.begin(), tmp.end()) - tmp.begin();
        }
    }
    for (int64_t j = 0; j < n; ++j) ucolptr[j + 1] += ucolptr[j];
    const int64_t nnz = ucolptr[n];
    NEPB_CHECK_ARG(nnz < (int64_t)std::numeric_limits<int32_t>::max() && n < (int64_t)std::numeric_limits<int32_t>::max(),
                   "int32 overflow: n=%lld nnz_union=%lld", (long long)n, (long long)nnz);

    // ---- pass 2: union row indices (CSC) -------------------------------------------------------
    u.n = n; u.p = p; u.nnz = nnz; u.vw = vw;
    u.colptr.assign(ucolptr.begin(), ucolptr.end());
    u.rowval.resize(nnz);
#pragma omp parallel
    {
        std::vector<int64_t> tmp;
#pragma omp for schedule(static)
        for (int64_t j = 0; j < n; ++j) {
            tmp.clear();
            for (int i = 0; i < p; ++i)
                for (int64_t e = colptr[i][j] - base; e < colptr[i][j + 1] - base; ++e) tmp.push_back(rowval[i][e] - base);
            std::sort(tmp.begin(), tmp.end());
            size_t m = std::unique(tmp.begin(), tmp.end()) - tmp.begin();
            int32_t* dst = u.rowval.data() + ucolptr[j];
            for (size_t t = 0; t < m; ++t) dst[t] = (int32_t)tmp[t];
        }
    }

    // ---- CSC -> CSR (counting sort by row; columns come out ascending within a row) ------------
    u.rowptr.assign(n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) u.rowptr[u.rowval[e] + 1]++;
    for (int64_t r = 0; r < n; ++r) u.rowptr[r + 1] += u.rowptr[r];
    u.colind.resize(nnz);
    u.csr_of_csc.resize(nnz);
    {
        std::vector<int32_t> next(u.rowptr.begin(), u.rowptr.end() - 1);
        for (int64_t j = 0; j < n; ++j)
            for (int64_t e = ucolptr[j]; e < ucolptr[j + 1]; ++e) {
                int32_t t = next[u.rowval[e]]++;
                u.colind[t] = (int32_t)j;
                u.csr_of_csc[e] = t;
            }
    }

    // ---- values, interleaved per union nonzero, CSR order --------------------------------------
    u.vals.assign((size_t)nnz * vw, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) {
        const int32_t* urow = u.rowval.data() + ucolptr[j];
        const int64_t ulen = ucolptr[j + 1] - ucolptr[j];
        for (int i = 0; i < p; ++i) {
            const double* src = (const double*)nzval[i];
            int64_t t = 0;
            for (int64_t e = colptr[i][j] - base; e < colptr[i][j + 1] - base; ++e) {
                int32_t r = (int32_t)(rowval[i][e] - base);
                while (t < ulen && urow[t] < r) ++t;  // both ascending: linear merge
                size_t dst = (size_t)u.csr_of_csc[ucolptr[j] + t] * vw + (size_t)i * vs;
                u.vals[dst] = src[e * vs];
                if (vs == 2) u.vals[dst + 1] = src[e * vs + 1];
            }
        }
    }
    return NEPB_OK;
}

}  // namespace nepb
