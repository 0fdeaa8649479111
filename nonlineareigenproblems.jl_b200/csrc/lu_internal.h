// Internal layouts shared by lu.cu and contour.cu.
#pragma once
#include <string.h>

#include <map>
#include <utility>
#include <vector>

#include "common.h"
#include "lu_symbolic.h"

namespace nepb {

struct LuDev {  // passed by value to kernels
    const int64_t* front_off;
    const int32_t* nf;
    const int32_t* np;
    const int32_t* ld;         // leading dimension of the front storage
    const uint8_t* in_place;   // the front's contribution block is its parent's front
    const uint8_t* has_ip;     // the front has such a child
    const int32_t* bw_slot;    // first partial-sum slot of a big front in the backward solve
    const int32_t* xsplit;     // update rows [0, xsplit) are pivot columns of the parent
    const int32_t* kback;      // delayed Schur updates: pivot columns of the earlier links of the window (0: none)
    const int64_t* row_ptr;
    const int32_t* rows;
    const int64_t* rel_ptr;
    const int32_t* rel;
    const int32_t* sn_ptr;
    const int64_t* w_off;
    const int32_t* child_ptr;
    const int32_t* child_list;
    int64_t front_total, w_total;
    int n;
};

struct EaRec {  // one child of one extend-add item (32 bytes)
    int64_t child_off;  // front offset of the child
    int64_t rel_off;    // its relative indices in the parent
    int32_t ldc, npc, ncb;
    int32_t xa, xb;     // child contribution-block columns landing in the item's slab
    int32_t pad;
};

struct LuInfo {
    unsigned long long amax_bits;    // max |M_ij| of the matrix that is factorised (bits of a non-negative double)
    unsigned long long amax_plain_bits;  // max |M_ij| of the operator itself (differs under static-pivoting scaling)
    unsigned long long minpiv_bits;  // min |pivot| / amax
    int flags;                       // 1: exactly zero pivot met (perturbed), 2: non-finite pivot
    int nperturbed;
};

// zero / non-finite / lifted pivots, or a pivot below 1e-8 max|M_ij|: with pivoting restricted to the pivot block this is how
// an unsafe elimination order shows (huge multipliers follow); it triggers the static-pivoting fallback on the plain pattern
// and a backward-error check of every solve
inline bool lu_info_suspicious(const LuInfo& I) {
    double r;
    memcpy(&r, &I.minpiv_bits, 8);
    return (I.flags & 3) || I.nperturbed > 0 || !(r >= 1e-8);
}

struct LuLevel {
    int front_begin = 0, front_count = 0, max_np = 1;
    int ea_begin = 0, ea_count = 0;
    int pn_begin = 0, pn_count = 0;
    int sc_begin = 0, sc_count = 0;
    int sp_begin = 0, sp_count = 0;  // pipelined schur items
    int fu_begin = 0, fu_count = 0, fu_count_all = 0;  // forward update items (big fronts); _all: including chain tails
    int bp_begin = 0, bp_count = 0;  // backward partial-product items (big fronts)
    int sfr_begin = 0, sfr_count = 0, sfr_count_all = 0;  // fronts handled per level in the solves (chain tails excluded / included)
    int fc_begin = 0, fc_count = 0;    // chains whose tail starts at this level (forward)
    int bc_begin = 0, bc_count = 0;    // chains whose tail ends at this level (backward)
};

struct LuSymbolicDev {
    LuSymbolic S;
    LuDev dev;
    std::vector<LuLevel> lv;
    DevBuf<int64_t> front_off, row_ptr, rel_ptr, w_off, a_pos;
    DevBuf<uint8_t> in_place, has_ip;
    DevBuf<int32_t> kback;
    bool schur_ring = false;
    int schur_tn = 64;  // column-tile step of the Schur items
    DevBuf<int32_t> bw_slot, xsplit, sfr_items, chain_fronts;
    DevBuf<int2> fc_items, bc_items;
    int part_slots = 0;
    DevBuf<int32_t> nf, np, ld, rows, rel, sn_ptr, child_ptr, child_list, perm, iperm, rperm, fr_items;
    DevBuf<double> dr, dc, a_scale;  // static pivoting (empty = none): row / column scalings, per-nonzero product
    bool matched() const { return a_scale.p != nullptr; }
    DevBuf<int4> ea_items, pn_items, sc_items, sp_items, fu_items, bp_items;
    DevBuf<EaRec> ea_recs;
    size_t smem_diag = 0, smem_panel = 0, smem_schur = 0, smem_schur_pipe = 0;
    bool schur_pipe = false;
    size_t smem_schur_dmma = 0;
    bool schur_dmma = false;
};

int lu_symbolic_get(const nepb_spmf* h, LuSymbolicDev** out);
int lu_symbolic_make_matched(const nepb_spmf* h, const double* coef, LuSymbolicDev** out);

}  // namespace nepb

struct nepb_lu {
    const nepb_spmf* op = nullptr;
    nepb::LuSymbolicDev* sym = nullptr;
    int nb = 0;   // shifts currently factorised
    int cap = 0;  // shifts the buffers can hold
    nepb::DevBuf<double> fronts;  // [nb][front_total] complex
    nepb::DevBuf<int> piv;        // [nb][n] local pivot rows
    nepb::DevBuf<nepb::LuInfo> info;
    nepb::DevBuf<double> coef;    // [nb][p] complex
    std::vector<double> h_coef;
    std::vector<nepb::LuInfo> h_info;
    // scratch
    nepb::DevBuf<double> xp, w, part, rhs, sol, res, cor, stage;
    nepb::DevBuf<unsigned long long> colmax;
    // captured single-shift solves, keyed by (shift, right-hand sides); valid while the staging buffers keep their address
    struct SolveGraph {
        cudaGraphExec_t exec = nullptr;
        int kernels = 0;
        const double* rhs = nullptr;
        const double* sol = nullptr;
    };
    std::map<std::pair<int, int>, SolveGraph> solve_graphs;
    ~nepb_lu() {
        for (auto& kv : solve_graphs) cudaGraphExecDestroy(kv.second.exec);
    }
};

namespace nepb {
int lu_create(const nepb_spmf* h, int nshift, const double* coef, nepb_lu** out);
// factorise shifts already described by lu->coef (device); used by the contour loop to recycle the storage
int lu_refactor(nepb_lu* lu, int nshift, const double* coef);
int lu_factor_device(nepb_lu* lu);                 // device work only (capturable); coefficients already in lu->coef
int lu_solve_reserve(nepb_lu* lu, int nb, int k);  // grow scratch outside capture
// solve for shifts [shift0, shift0+nb): Bdev [b][n][k] (rhs_stride = n*k) or shared (0); Xdev [b][n][k]
int lu_solve_device(nepb_lu* lu, int shift0, int nb, int k, const double2* Bdev, size_t rhs_stride, double2* Xdev);
// factorise lu->nb shifts and solve them in one pipelined sequence: forward substitution of level l on `side` beside the
// factorisation of the levels above it; ev = 3 * nlevels + 2 events
int lu_factor_solve_pipelined(nepb_lu* lu, int k, const double2* Bdev, size_t rhs_stride, double2* Xdev, cudaStream_t side,
                              cudaEvent_t* ev);
}  // namespace nepb
