// Host-side symbolic analysis for the device multifrontal LU (one per SPMF union pattern).
#pragma once
#include <stdint.h>
#include <vector>

namespace nepb {

struct LuSymbolic {
    int n = 0;
    int64_t nnz = 0;
    // ordering: perm[new] = old, iperm[old] = new (fill-reducing ordering composed with the etree postorder)
    std::vector<int32_t> perm, iperm;
    std::vector<int32_t> rowmap;  // empty, or row i of the operator = row rowmap[i] of the factorised matrix (static pivoting)
    std::vector<int32_t> parent;  // column elimination tree of the permuted pattern of A + A^T (-1 = root)
    std::vector<int32_t> colcount;  // |struct(L_j)| including the diagonal
    // supernodes = fronts
    int nsuper = 0;
    std::vector<int32_t> sn_ptr;     // [nsuper+1] first pivot column of each supernode
    std::vector<int32_t> sn_parent;  // supernodal tree (-1 = root)
    std::vector<int32_t> col_sn;     // [n] supernode of a permuted column
    std::vector<int64_t> row_ptr;    // [nsuper+1] offsets into rows
    std::vector<int32_t> rows;       // front row structure: np pivots first, then the update rows (ascending)
    std::vector<int64_t> front_off;  // [nsuper] offsets (in complex elements) of the fronts; [nsuper] = total storage
    std::vector<int32_t> front_ld;   // [nsuper] leading dimension (nf, or the owner's when the front lives inside its child)
    std::vector<uint8_t> in_place_child;  // [nsuper] 1 = this front's contribution block IS its parent's front
    std::vector<uint8_t> has_in_place_child;  // [nsuper] 1 = one of the children is in place
    std::vector<int64_t> rel_ptr;    // [nsuper+1] offsets into rel
    std::vector<int32_t> rel;        // for the update rows of s: their position in the parent's row structure
    std::vector<int32_t> level;      // [nsuper] height above the leaves
    int nlevels = 0;
    std::vector<int32_t> level_ptr, level_list;  // supernodes grouped by level
    std::vector<int64_t> a_pos;      // [nnz] CSR nonzero -> offset in the front storage
    std::vector<int64_t> w_off;      // [nsuper] offsets into the solve work rows (nf each; in-place parents alias); [nsuper] = total
    int64_t front_total = 0, nnz_factor = 0, w_total = 0;
    double flops = 0;  // complex multiply-adds of the numeric factorisation
    int max_nf = 0, max_np = 0;
};

struct LuOptions {
    int relax_leaf = 32;   // subtrees with at most this many columns become one supernode (gun: 32 beats 16 and 8 on the B200)
    int max_np = 32;       // cap on pivot columns per front (wider supernodes are split into chains)
    int ordering = 0;      // 0 = approximate minimum degree on A + A^T, 1 = natural
    int alias_chains = 1;  // parent fronts with one structurally identical child live inside that child's storage
};

// csr rowptr/colind of the n x n union pattern (0-based, int32); user_perm optional (perm[new] = old)
int lu_symbolic_analyse(int n, const int32_t* rowptr, const int32_t* colind, const int32_t* user_perm,
                        const LuOptions& opt, LuSymbolic& S, const int32_t* rowmap = nullptr);

// maximum-product matching of rows to columns with I-matrix scalings (lu_matching.cpp); returns the matched rows (n = success)
int max_product_matching(int n, const int32_t* rowptr, const int32_t* colind, const double* absval,
                         std::vector<int32_t>& row_to_col, std::vector<double>& dr, std::vector<double>& dc);

// approximate-minimum-degree ordering of the symmetric graph (adjacency without diagonal); out[new] = old
void amd_order(int n, const std::vector<int64_t>& xadj, const std::vector<int32_t>& adj, std::vector<int32_t>& out);

}  // namespace nepb
