// Device multifrontal LU of M(sigma) = sum_i c_i A_i on the SPMF union pattern, batched over shifts, and the
// batched multi-RHS triangular solves against it (sm_100a).
//
// Replaces (reference, relative to src/): LinSolvers.jl:109-137 (FactorizeLinSolver: factorize(compute_Mder) +
// `Afact \ x`), :147-159 (BackslashLinSolver), LinSolverCreators.jl:62-122; the N factor+solve pairs of
// method_contour_common.jl:81-91.  The arithmetic the reference delegates to UMFPACK is restated here as a
// multifrontal method because that is what maps onto the GPU: the symbolic analysis (lu_symbolic.cpp) is done once
// per sparsity pattern, every shift re-uses it, and a *batch* of shifts is factorised concurrently -- the batch
// and the independent fronts of one elimination-tree level are the two sources of parallelism.
//
// Per level of the supernodal tree:
//   extend-add   parent fronts gather their children's contribution blocks (parent-centric, no atomics,
//                deterministic summation order)
//   diag         partial-pivoted LU of the np x np pivot block in shared memory (pivoting restricted to the
//                fully-summed rows of the front; tiny pivots are perturbed and counted, zero / non-finite
//                pivots flag the shift as singular)
//   panel        U12 = L11^-1 P F12 and L21 = F21 U11^-1, tiled over CTAs
//   schur        F22 -= L21 U12, 64 x 64 output tiles (FP64 FMA pipe; B200 has no faster FP64 path)
// Fronts are dense column-major nf x nf blocks, all fronts of one shift contiguous (stride front_total).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.h"
#include "lu_symbolic.h"
#include "lu_internal.h"

namespace nepb {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void cfms(double2& acc, const double2 a, const double2 b) {  // acc -= a*b
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfma2(double2& acc, const double2 a, const double2 b) {  // acc += a*b
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// 1/x from the hardware approximation + two Newton steps (full double accuracy for normal x; IEEE division costs an
// order of magnitude more instructions and sits on the serial path of the pivot-block factorisation)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double2 crecip(double2 a) {
    // 1/a = conj(a)/|a|^2 with the operand scaled by a power of two first: no overflow / underflow of |a|^2
    const double m = fmax(fabs(a.x), fabs(a.y));
    const int e = (int)((__double2hiint(m) >> 20) & 0x7ff) - 1023;              // exponent of the larger component
    const double sc = __hiloint2double((1023 - e) << 20, 0);                    // 2^-e (e in the normal range)
    const double xs = a.x * sc, ys = a.y * sc;
    const double r = fast_rcp(fma(xs, xs, ys * ys)) * sc;
    return make_double2(xs * r, -ys * r);
}
__device__ __forceinline__ double cabs1(double2 a) { return fabs(a.x) + fabs(a.y); }

// ---------------------------------------------------------------------------------------------
// assembly: fronts[b][a_pos[e]] = sum_i coef[b][i] * vals[e][i]   (fused compute_Mder + scatter)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lu_assemble_kernel(int64_t nnz, int p, int ca, const int64_t* __restrict__ a_pos,
                                                          const double* __restrict__ a_scale, const double* __restrict__ vals,
                                                          const double2* __restrict__ coef, double2* __restrict__ fronts,
                                                          int64_t front_total, LuInfo* __restrict__ info) {
    const int b = blockIdx.y;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double2* c = coef + (size_t)b * p;
    double a = 0.0, a0 = 0.0;
    if (e < nnz) {
        const int vw = ca ? 2 * p : p;
        const double* v = vals + (size_t)e * vw;
        double2 m = make_double2(0.0, 0.0);
        for (int i = 0; i < p; ++i) {
            const double2 x = ca ? make_double2(v[2 * i], v[2 * i + 1]) : make_double2(v[i], 0.0);
            cfma2(m, c[i], x);
        }
        a0 = cabs1(m);
        if (a_scale) {  // static pivoting: D_r M D_c (lu_matching.cpp)
            const double sc = a_scale[e];
            m.x *= sc;
            m.y *= sc;
        }
        fronts[(size_t)b * front_total + a_pos[e]] = m;
        a = cabs1(m);
    }
    // max |entry| of M(sigma_b): scale for the tiny-pivot test
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a = fmax(a, __shfl_xor_sync(0xffffffffu, a, off));
        a0 = fmax(a0, __shfl_xor_sync(0xffffffffu, a0, off));
    }
    if ((threadIdx.x & 31) == 0 && a0 > 0.0 && a_scale) atomicMax((unsigned long long*)&info[b].amax_plain_bits, (unsigned long long)__double_as_longlong(a0));
    if ((threadIdx.x & 31) == 0 && a > 0.0) {
        atomicMax((unsigned long long*)&info[b].amax_bits, (unsigned long long)__double_as_longlong(a));
        if (!a_scale) atomicMax((unsigned long long*)&info[b].amax_plain_bits, (unsigned long long)__double_as_longlong(a));
    }
}

// ---------------------------------------------------------------------------------------------
// extend-add: item = (parent s, -, first record, number of records) for one column slab [j0, j1) of the parent.  A record
// describes one child whose contribution block has columns landing in the slab: the column range [xa, xb) (found on the
// host: rel is strictly increasing) and everything the copy needs, so the kernel makes no dependent look-ups.
// Children are applied one after the other (fixed summation order); within a child every parent entry is hit once.
// A warp walks one child column at a time, lanes along its rows: coalesced reads of the contribution block.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lu_extend_add_kernel(LuDev d, const int4* __restrict__ items, const EaRec* __restrict__ recs,
                                                            double2* __restrict__ fronts) {
    const int4 it = items[blockIdx.x];
    const int s = it.x;
    const int b = blockIdx.y;
    double2* base = fronts + (size_t)b * d.front_total;
    const int ldp = d.ld[s];
    double2* Fp = base + d.front_off[s];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int q = 0; q < it.w; ++q) {
        const EaRec r = recs[it.z + q];
        const int* __restrict__ rel = d.rel + r.rel_off;
        const double2* __restrict__ Fc = base + r.child_off + (size_t)r.npc * (r.ldc + 1);  // first entry of the contribution block
        for (int x = r.xa + wid; x < r.xb; x += 8) {
            const double2* col = Fc + (size_t)x * r.ldc;
            double2* pcol = Fp + (size_t)rel[x] * ldp;
            for (int y = lane; y < r.ncb; y += 32) {
                const double2 v = col[y];
                double2* t = pcol + rel[y];
                double2 o = *t;
                o.x += v.x;
                o.y += v.y;
                *t = o;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// diag: partial-pivoted LU of the (at most 32 x 32) pivot block AND the explicit inverses of its triangular factors.
// One CTA of 4 warps per (front, shift).  Both phases are short rolled loops (the kernel sits on the critical path of
// every tree level, and straight-line code of this size is bound by instruction fetch, not by arithmetic).
//   LU phase   lane = physical row (rows are never exchanged: a row remembers the step at which it became the pivot
//              row); warp w owns the columns w, w+4, ... in a register window whose first entry is always its next
//              column to eliminate (the window is rotated after use, so all register indices are static).  Per step:
//              the owner warp picks the pivot with one redux.max over float-rounded magnitudes + ballot (threshold 0.1
//              in favour of row j), forms the multipliers and publishes them through double-buffered shared memory;
//              one block barrier; every warp applies the rank-1 update to its remaining columns with the pivot row
//              broadcast by shuffles.
//   inverses   warp 0: inv(L11), warp 1: inv(U11); lane = column of the inverse, right-looking substitution on a
//              rotating register window; factor entries are shared-memory broadcasts.  No inter-lane dependency.
// The inverses turn every later triangular solve with this block (panel, forward, backward) into a small dense product.
// Output in place of F11: strict lower part = inv(L11) (unit diagonal implied), upper part = inv(U11);
// piv[t] = original row (within the block) that became pivot row t.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 shfl_c(double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

constexpr int DNP = 32, DCW = 8, DNW = 4;
constexpr int DLD = 3 * DNP + 1;  // row of the factor block in shared memory: 32 zeros | 32 entries | 32 zeros (+1 pad)
constexpr size_t DIAG_SMEM = ((size_t)DNP * DLD + 2 * DNP + DNP) * 16 + 16;

// column `lane` of inv(L11) (UP = false) or inv(U11) (UP = true) by right-looking substitution on a register window:
// y[m] = x[t + m] resp. x[t - m], so y[0] is always the entry being finished and every register index is static.
// Rows of the factor block are zero-padded on both sides, so the window needs no bounds checks.  W = window length >= np.
template <int W, bool UP>
__device__ __forceinline__ void diag_tri_inverse(const double2* __restrict__ sLU, const double2* __restrict__ s_rp, int np, int lane,
                                                 double2* __restrict__ F, int ld) {
    double2 y[W];
#pragma unroll
    for (int m = 0; m < W; ++m) y[m] = make_double2((UP ? np - 1 - m : m) == lane ? 1.0 : 0.0, 0.0);
#pragma unroll 1
    for (int q = 0; q < np; ++q) {
        const int t = UP ? np - 1 - q : q;
        double2 xt = y[0];
        if (UP) xt = cmul(xt, s_rp[t]);
        if (lane < np && (UP ? t <= lane : t > lane)) F[(size_t)t + (size_t)lane * ld] = xt;
        const double2* row = sLU + t * DLD + DNP + t;
#pragma unroll
        for (int m = 1; m < W; ++m) {
            const double2 f = UP ? row[-m] : row[m];
            double2 r = y[m];
            cfms(r, f, xt);
            y[m - 1] = r;
        }
        y[W - 1] = make_double2(0.0, 0.0);
    }
}

__global__ void __launch_bounds__(128) lu_diag_inv_kernel(LuDev d, const int* __restrict__ items, double2* __restrict__ fronts,
                                                          int* __restrict__ piv, LuInfo* __restrict__ info) {
    extern __shared__ double2 dsm[];
    double2* sLU = dsm;                      // [c][DLD]: sLU[c*DLD + DNP + r] = (L\U)[r][c]
    double2* s_l = dsm + DNP * DLD;          // [2][DNP] multipliers of the current step
    double2* s_rp = s_l + 2 * DNP;           // [DNP] reciprocal pivots
    int* s_p = (int*)(s_rp + DNP);           // [2] pivot row of the current step
    const int s = items[blockIdx.x];
    const int b = blockIdx.y;
    const int np = d.np[s], ld = d.ld[s];
    double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    int* pv = piv + (size_t)b * d.n + d.sn_ptr[s];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double2 a[DCW];  // window over the columns w, w+4, ..: a[0] is the next column this warp eliminates
#pragma unroll
    for (int c = 0; c < DCW; ++c) {
        const int cg = DNW * c + w;
        a[c] = (lane < np && cg < np) ? F[(size_t)lane + (size_t)cg * ld] : make_double2(lane == cg ? 1.0 : 0.0, 0.0);
    }
    for (int idx = threadIdx.x; idx < np * (2 * DNP + 1); idx += 128) {  // zero padding of the rows that will be read
        const int r = idx / (2 * DNP + 1), x = idx % (2 * DNP + 1);
        sLU[r * DLD + (x < DNP ? x : x + DNP)] = make_double2(0.0, 0.0);
    }
    if (w == 0) s_rp[lane] = make_double2(1.0, 0.0);
    const double amax = __longlong_as_double(info[b].amax_bits);
    const double tiny = 2.220446049250313e-16 * amax;
    double minpiv = INFINITY;
    int nperturbed = 0, flags = 0;
    int pos = (lane < np) ? -1 : lane;  // step at which this row became the pivot row
    // the owner warp of column j (its window starts at that column, fully updated) picks the pivot row, forms the multipliers
    // and publishes them; then it rotates its window
    auto owner_step = [&](int j) {
        const int par = j & 1;
        double2 aj = a[0];
        const double mine = cabs1(aj);
        const bool cand = pos < 0;
        const float mf = (mine == mine) ? (float)mine : INFINITY;
        const unsigned key = cand ? __float_as_uint(mf) + 1u : 0u;  // non-negative floats order like their bit patterns
        const unsigned best = __reduce_max_sync(0xffffffffu, key);
        int bi = __ffs(__ballot_sync(0xffffffffu, key == best)) - 1;
        const unsigned kj = __shfl_sync(0xffffffffu, key, j);
        if (kj != 0u && __uint_as_float(kj - 1u) >= 0.1f * __uint_as_float(best - 1u)) bi = j;  // prefer the diagonal
        double2 pvt = shfl_c(aj, bi);
        const double pa = cabs1(pvt);
        if (!(pa <= 1.79e308)) {
            flags |= 2;
            pvt = make_double2(1.0, 0.0);
        } else if (pa == 0.0) {
            flags |= 1;
            ++nperturbed;
            pvt = make_double2(tiny > 0.0 ? tiny : 1.0, 0.0);
        } else if (pa < tiny) {
            ++nperturbed;
            const double sc = tiny / pa;
            pvt = make_double2(pvt.x * sc, pvt.y * sc);
        }
        minpiv = fmin(minpiv, pa);
        const double2 rp = crecip(pvt);
        double2 l = make_double2(0.0, 0.0);
        if (lane == bi) {
            aj = pvt;
            pos = j;
        } else if (cand) {
            l = cmul(aj, rp);
            aj = l;
        }
        sLU[j * DLD + DNP + lane] = aj;  // column j is final: multipliers below the pivot, U entries in the rows chosen earlier
        s_l[par * DNP + lane] = l;
        if (lane == 0) {
            s_p[par] = bi;
            s_rp[j] = rp;
        }
#pragma unroll
        for (int c = 0; c < DCW - 1; ++c) a[c] = a[c + 1];
    };
    __syncthreads();  // padding written before any column lands
    if (w == 0) owner_step(0);
#pragma unroll 1
    for (int j = 0; j < np; ++j) {
        const int par = j & 1, owner = j & (DNW - 1);
        __syncthreads();
        const int p = s_p[par];
        const double2 l = s_l[par * DNP + lane];  // zero for the pivot row and for rows that became pivot rows earlier
        if (w != owner && lane == p) pos = j;
        const int first = (j & ~(DNW - 1)) + w + (w > owner ? 0 : DNW);  // first column of this warp beyond j
        int rem = (np - first + DNW - 1) / DNW;                          // its columns below np
        if (j + 1 < np && w == ((j + 1) & (DNW - 1))) {
            // look-ahead: the owner of the next column updates that column first and eliminates it right away, so the
            // serial chain per step is one column update + the pivot search, not the whole rank-1 update
            const double2 rj = shfl_c(a[0], p);
            cfms(a[0], l, rj);
            owner_step(j + 1);
            --rem;
        }
#pragma unroll
        for (int c = 0; c < DCW; ++c) {
            if (c < rem) {  // warp-uniform
                const double2 rj = shfl_c(a[c], p);
                cfms(a[c], l, rj);
            }
        }
    }
    __syncthreads();
    {  // rows into pivot order
        double2 tmp[DCW];
#pragma unroll
        for (int c = 0; c < DCW; ++c) tmp[c] = (DNW * c + w < np) ? sLU[(DNW * c + w) * DLD + DNP + lane] : make_double2(0.0, 0.0);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < DCW; ++c)
            if (DNW * c + w < np) sLU[(DNW * c + w) * DLD + DNP + pos] = tmp[c];
        if (w == 0 && lane < np) pv[pos] = lane;
        __syncthreads();
    }
    if (np <= 16) {
        if (w == 0) diag_tri_inverse<16, false>(sLU, s_rp, np, lane, F, ld);
        else if (w == 1) diag_tri_inverse<16, true>(sLU, s_rp, np, lane, F, ld);
    } else {
        if (w == 0) diag_tri_inverse<DNP, false>(sLU, s_rp, np, lane, F, ld);
        else if (w == 1) diag_tri_inverse<DNP, true>(sLU, s_rp, np, lane, F, ld);
    }
    if (lane == 0) {
        if (flags) atomicOr(&info[b].flags, flags);
        if (nperturbed) atomicAdd(&info[b].nperturbed, nperturbed);
        if (minpiv < INFINITY) {
            const double ratio = amax > 0.0 ? minpiv / amax : 0.0;
            atomicMin((unsigned long long*)&info[b].minpiv_bits, (unsigned long long)__double_as_longlong(ratio));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// panel: item = (front s, kind 0 = U tile / 1 = L tile, tile start t0), 64 columns (rows) per tile, as dense products
// with the inverted pivot block:  U12 = inv(L11) (P F12),  L21 = F21 inv(U11).
// ---------------------------------------------------------------------------------------------
constexpr int PANEL_T = 64;
constexpr size_t PANEL_SMEM = ((size_t)DNP * (DNP + 1) + (size_t)DNP * (PANEL_T + 1)) * 16;
__global__ void __launch_bounds__(256) lu_panel_inv_kernel(LuDev d, const int4* __restrict__ items, double2* __restrict__ fronts,
                                                           const int* __restrict__ piv) {
    extern __shared__ double2 sm[];
    double2* sI = sm;                    // [t][i], stride DNP+1: inv(L11) (kind 0, unit diagonal) or inv(U11) (kind 1)
    double2* sB = sm + DNP * (DNP + 1);  // kind 0: [t][c] stride PANEL_T+1 ; kind 1: [t][r] stride PANEL_T
    const int4 it = items[blockIdx.x];
    const int s = it.x, kind = it.y, t0 = it.z;
    const int b = blockIdx.y;
    const int ld = d.ld[s], np = d.np[s], ncb = d.nf[s] - np;
    double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    const int* pv = piv + (size_t)b * d.n + d.sn_ptr[s];
    const int tid = threadIdx.x;
    const int tw = min(PANEL_T, ncb - t0);
    if (kind == 0) {
        // sI[t][i] = inv(L11)[i][t]
        for (int idx = tid; idx < DNP * DNP; idx += 256) {
            const int i = idx % DNP, t = idx / DNP;
            double2 v = make_double2(i == t ? 1.0 : 0.0, 0.0);
            if (i < np && t < i) v = F[(size_t)i + (size_t)t * ld];
            sI[t * (DNP + 1) + i] = v;
        }
        for (int idx = tid; idx < np * PANEL_T; idx += 256) {
            const int t = idx % np, c = idx / np;
            sB[t * (PANEL_T + 1) + c] = (c < tw) ? F[(size_t)pv[t] + (size_t)(np + t0 + c) * ld] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        const int i = tid & 31, cs = tid >> 5;
        double2 acc[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) acc[m] = make_double2(0.0, 0.0);
        for (int t = 0; t < np; ++t) {
            const double2 l = sI[t * (DNP + 1) + i];
            const double2* bb = sB + t * (PANEL_T + 1) + cs;
#pragma unroll
            for (int m = 0; m < 8; ++m) cfma2(acc[m], l, bb[8 * m]);
        }
        if (i < np) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int c = cs + 8 * m;
                if (c < tw) F[(size_t)i + (size_t)(np + t0 + c) * ld] = acc[m];
            }
        }
    } else {
        // sI[t][c] = inv(U11)[t][c] (zero below the diagonal)
        for (int idx = tid; idx < DNP * DNP; idx += 256) {
            const int t = idx % DNP, c = idx / DNP;
            double2 v = make_double2(0.0, 0.0);
            if (c < np && t <= c) v = F[(size_t)t + (size_t)c * ld];
            sI[t * (DNP + 1) + c] = v;
        }
        for (int idx = tid; idx < np * PANEL_T; idx += 256) {
            const int r = idx % PANEL_T, t = idx / PANEL_T;
            sB[t * PANEL_T + r] = (r < tw) ? F[(size_t)(np + t0 + r) + (size_t)t * ld] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        const int r = tid & 63, cg = (tid >> 6) * 8;
        double2 acc[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) acc[m] = make_double2(0.0, 0.0);
        for (int t = 0; t < np; ++t) {
            const double2 f = sB[t * PANEL_T + r];
            const double2* uu = sI + t * (DNP + 1) + cg;
#pragma unroll
            for (int m = 0; m < 8; ++m) cfma2(acc[m], f, uu[m]);
        }
        if (r < tw) {
#pragma unroll
            for (int m = 0; m < 8; ++m)
                if (cg + m < np) F[(size_t)(np + t0 + r) + (size_t)(cg + m) * ld] = acc[m];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// schur: item = (front s, row tile i0, column tile j0); F22[i0.., j0..] -= L21[i0.., :] * U12[:, j0..]
// 64 x 64 tile, 256 threads, each 4 x 4 outputs (rows tx + 16 r, columns ty + 16 c)
// ---------------------------------------------------------------------------------------------
constexpr int SCHUR_T = 64;
__global__ void __launch_bounds__(256) lu_schur_kernel(LuDev d, const int4* __restrict__ items, double2* __restrict__ fronts) {
    extern __shared__ double2 sm[];
    const int4 it = items[blockIdx.x];
    const int s = it.x, i0 = it.y, j0 = it.z;
    const int b = blockIdx.y;
    const int nf = d.ld[s], np = d.np[s], ncb = d.nf[s] - np;  // nf: leading dimension of the front storage
    double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    double2* sL = sm;                           // [np][SCHUR_T]   L21 tile, row index fastest
    double2* sU = sm + (size_t)np * SCHUR_T;    // [np][SCHUR_T+1] U12 tile transposed: sU[t][c] (padded: conflict-free transposing store)
    const int tid = threadIdx.x;
    const int th = min(SCHUR_T, ncb - i0), tw = min(SCHUR_T, ncb - j0);
    for (int idx = tid; idx < np * SCHUR_T; idx += blockDim.x) {
        const int r = idx % SCHUR_T, t = idx / SCHUR_T;
        sL[idx] = (r < th) ? F[(size_t)(np + i0 + r) + (size_t)t * nf] : make_double2(0.0, 0.0);
    }
    for (int idx = tid; idx < np * SCHUR_T; idx += blockDim.x) {
        const int t = idx % np, c = idx / np;  // contiguous reads along t
        sU[t * (SCHUR_T + 1) + c] = (c < tw) ? F[(size_t)t + (size_t)(np + j0 + c) * nf] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int tx = tid % 16, ty = tid / 16;
    double2 acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = make_double2(0.0, 0.0);
    for (int t = 0; t < np; ++t) {
        double2 l[4], u[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) l[r] = sL[t * SCHUR_T + tx + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) u[c] = sU[t * (SCHUR_T + 1) + ty + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) cfma2(acc[r][c], l[r], u[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int cc = ty + 16 * c;
        if (cc >= tw) continue;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rr = tx + 16 * r;
            if (rr >= th) continue;
            double2* t = F + (size_t)(np + i0 + rr) + (size_t)(np + j0 + cc) * nf;
            double2 o = *t;
            o.x -= acc[r][c].x;
            o.y -= acc[r][c].y;
            *t = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// schur, pipelined: item = (front s, row tile i0, first column tile j0, number of column tiles).  The L21 tile stays in
// shared memory for all column tiles of the item; U12 tiles are double-buffered with cp.async so that the next tile
// streams in while the current one is multiplied.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// C -= v as a fire-and-forget reduction performed in L2: no load, so the epilogue never waits on HBM.  Every element of
// the trailing matrix receives exactly one such update per level, so the result does not depend on any ordering.
__device__ __forceinline__ void red_sub_c(double2* addr, double2 v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&addr->x), "d"(-v.x) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&addr->y), "d"(-v.y) : "memory");
}

constexpr int SCHUR_GROUP = 2;  // column tiles per item
__global__ void __launch_bounds__(256, 2) lu_schur_pipe_kernel(LuDev d, const int4* __restrict__ items, double2* __restrict__ fronts) {
    extern __shared__ double2 sm[];
    const int4 it = items[blockIdx.x];
    const int s = it.x, i0 = it.y, jfirst = it.z, ntiles = it.w;
    const int b = blockIdx.y;
    const int nf = d.ld[s], np = d.np[s], ncb = d.nf[s] - np;
    double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    constexpr int ULD = SCHUR_T + 1;
    double2* sL = sm;                         // [np][SCHUR_T]
    double2* sU0 = sm + (size_t)np * SCHUR_T;  // 2 x [np][ULD]
    const int tid = threadIdx.x;
    const int th = min(SCHUR_T, ncb - i0);
    auto stage_u = [&](int j0, double2* sU) {
        const int tw = min(SCHUR_T, ncb - j0);
        for (int idx = tid; idx < np * SCHUR_T; idx += 256) {
            const int t = idx % np, c = idx / np;
            if (c < tw) cp_async16(&sU[t * ULD + c], &F[(size_t)t + (size_t)(np + j0 + c) * nf]);
            else sU[t * ULD + c] = make_double2(0.0, 0.0);
        }
    };
    for (int idx = tid; idx < np * SCHUR_T; idx += 256) {
        const int r = idx % SCHUR_T, t = idx / SCHUR_T;
        if (r < th) cp_async16(&sL[idx], &F[(size_t)(np + i0 + r) + (size_t)t * nf]);
        else sL[idx] = make_double2(0.0, 0.0);
    }
    stage_u(jfirst, sU0);
    cp_async_commit();
    const int tx = tid % 16, ty = tid / 16;
    for (int q = 0; q < ntiles; ++q) {
        const int j0 = jfirst + q * SCHUR_T;
        double2* sU = sU0 + (size_t)(q & 1) * np * ULD;
        if (q + 1 < ntiles) {
            stage_u(j0 + SCHUR_T, sU0 + (size_t)((q + 1) & 1) * np * ULD);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int tw = min(SCHUR_T, ncb - j0);
        double2 acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = make_double2(0.0, 0.0);
#pragma unroll 2
        for (int t = 0; t < np; ++t) {
            double2 l[4], u[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) l[r] = sL[t * SCHUR_T + tx + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; ++c) u[c] = sU[t * ULD + ty + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) cfma2(acc[r][c], l[r], u[c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int cc = ty + 16 * c;
            if (cc < tw) {
                double2* col = F + (size_t)(np + i0) + (size_t)(np + j0 + cc) * nf;
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (tx + 16 * r < th) red_sub_c(col + tx + 16 * r, acc[r][c]);
            }
        }
        __syncthreads();  // the buffer of tile q is refilled in iteration q + 1
    }
}

// ---------------------------------------------------------------------------------------------
// schur on the FP64 tensor pipe (round 2): same items as the pipelined kernel, the 64 x 64 x np product of a tile as
// mma.sync.m8n8k4.f64 (DMMA) -- 8 warps, each 16 rows x 32 columns = 2 x 4 MMA tiles, 4 real MMAs per complex product.
// The DFMA version reads 128 B of shared memory per thread and k-step for 64 DFMA: the 128 B/clk shared-memory pipe and the
// FP64 pipe are busy for the same number of cycles (measured: FP64 pipe 43 % active).  With MMA fragments a warp reads
// 12 x 256 B per 4 k-steps for 8192 DFMA: the shared-memory pipe needs a fifth of the FP64 time.
// Operands are de-interleaved into re / im planes by 8-byte cp.async copies (zero-fill outside the tile and for k >= np):
// L21 as [k][row] with a pitch of 68 doubles, U12 as [column][k] with a pitch of 36, so that both the copies (row-fastest /
// k-fastest, following the column-major fronts) and the fragment reads hit 16 different bank pairs per phase.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8z(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gsrc), "r"(sz));
}
__device__ int g_schur_dbg = 0;  // timing experiments (NEPB_LU_SCHUR_DBG): 1 no MMA, 2 no epilogue, 4 load / store epilogue, 8 no staging
constexpr int SD_LLD = SCHUR_T + 4;  // L planes: [k][row]
constexpr int SD_ULD = 36;           // U planes: [column][k], k <= 32
constexpr int SD_UBUF = 2 * SCHUR_T * SD_ULD;  // doubles per U buffer (re plane + im plane)
template <bool CINIT>
__global__ void __launch_bounds__(256, 2) lu_schur_dmma_kernel(LuDev d, const int4* __restrict__ items, double2* __restrict__ fronts) {
    extern __shared__ double sdm[];
    const int4 it = items[blockIdx.x];
    const int s = it.x, i0 = it.y, jfirst = it.z, ntiles = it.w;
    const int b = blockIdx.y;
    const int nf = d.ld[s], np = d.np[s], ncb = d.nf[s] - np;
    const int npad = (np + 3) & ~3;
    double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    double* sLr = sdm;
    double* sLi = sLr + npad * SD_LLD;
    double* sU0 = sLi + npad * SD_LLD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int th = min(SCHUR_T, ncb - i0);
    const int dbg = g_schur_dbg;
    auto stage_u = [&](int j0, double* sU) {
        const int tw = min(SCHUR_T, ncb - j0);
        if (dbg & 8) return;
        for (int idx = tid; idx < npad * SCHUR_T; idx += 256) {
            const int t = idx % npad, c = idx / npad;
            const bool ok = c < tw && t < np;
            const double* src = (const double*)(F + (ok ? (size_t)t + (size_t)(np + j0 + c) * nf : 0));
            cp_async8z(sU + c * SD_ULD + t, src, ok);
            cp_async8z(sU + SCHUR_T * SD_ULD + c * SD_ULD + t, src + 1, ok);
        }
    };
    for (int idx = tid; idx < ((dbg & 8) ? 0 : npad * SCHUR_T); idx += 256) {
        const int r = idx % SCHUR_T, t = idx / SCHUR_T;
        const bool ok = r < th && t < np;
        const double* src = (const double*)(F + (ok ? (size_t)(np + i0 + r) + (size_t)t * nf : 0));
        cp_async8z(sLr + t * SD_LLD + r, src, ok);
        cp_async8z(sLi + t * SD_LLD + r, src + 1, ok);
    }
    stage_u(jfirst, sU0);
    cp_async_commit();
    const int ar = lane >> 2, ak = lane & 3;
    const int wr = (warp & 3) * 16, wc = (warp >> 2) * 32;
    for (int q = 0; q < ntiles; ++q) {
        const int j0 = jfirst + q * SCHUR_T;
        const double* sUr = sU0 + (size_t)(q & 1) * SD_UBUF;
        const double* sUi = sUr + SCHUR_T * SD_ULD;
        const int tw = min(SCHUR_T, ncb - j0);
        double cr[2][4][2], ci[2][4][2];
        // CINIT: the accumulators start as the C tile itself (loads in flight while the operand tiles land), the product is
        // formed with -L21, and the tile is written back with plain stores: no atomics, the tile is owned by this CTA
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cc = wc + nt * 8 + 2 * ak + e;
                const double2* col = F + (size_t)(np + i0) + (size_t)(np + j0 + cc) * nf;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int rr = wr + mt * 8 + ar;
                    double2 v = make_double2(0.0, 0.0);
                    if (CINIT && cc < tw && rr < th) v = col[rr];
                    cr[mt][nt][e] = v.x;
                    ci[mt][nt][e] = v.y;
                }
            }
        if (q + 1 < ntiles) {
            stage_u(j0 + SCHUR_T, sU0 + (size_t)((q + 1) & 1) * SD_UBUF);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (wr < th && wc < tw) {  // warp-uniform: skip warps whose 16 x 32 patch lies outside the tile
#pragma unroll 2
            for (int k4 = 0; k4 < ((dbg & 1) ? 0 : npad); k4 += 4) {
                double a_r[2], a_i[2], na_i[2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    a_r[mt] = sLr[(k4 + ak) * SD_LLD + wr + mt * 8 + ar];
                    a_i[mt] = sLi[(k4 + ak) * SD_LLD + wr + mt * 8 + ar];
                    if (CINIT) {  // -L21
                        a_r[mt] = -a_r[mt];
                        a_i[mt] = -a_i[mt];
                    }
                    na_i[mt] = -a_i[mt];
                }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const double b_r = sUr[(wc + nt * 8 + ar) * SD_ULD + k4 + ak];
                    const double b_i = sUi[(wc + nt * 8 + ar) * SD_ULD + k4 + ak];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], a_r[mt], b_r);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], a_r[mt], b_i);
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], na_i[mt], b_i);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], a_i[mt], b_r);
                    }
                }
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int cc = wc + nt * 8 + 2 * ak + e;
                    if (cc < tw) {
                        double2* col = F + (size_t)(np + i0) + (size_t)(np + j0 + cc) * nf;
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const int rr = wr + mt * 8 + ar;
                            if (CINIT) {
                                if (rr < th) col[rr] = make_double2(cr[mt][nt][e], ci[mt][nt][e]);
                            } else if (rr < th && !(dbg & 2)) {
                                if (dbg & 4) {
                                    double2 o = col[rr];
                                    o.x -= cr[mt][nt][e];
                                    o.y -= ci[mt][nt][e];
                                    col[rr] = o;
                                } else {
                                    red_sub_c(col + rr, make_double2(cr[mt][nt][e], ci[mt][nt][e]));
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // the buffer of tile q is refilled in iteration q + 1
    }
}

// ---------------------------------------------------------------------------------------------
// schur with delayed updates (round 2, the default): the skip experiments of profiles/r2_contour_breakdown.txt show that the
// trailing update is bound by the read-modify-write of C in HBM (82 GB per 128-node step: every link of a supernode that was
// split at 32 pivot columns passes over the whole trailing matrix with an inner dimension of 32), not by the FP64 pipe.
// In a chain of in-place fronts the panels of consecutive links are neighbouring columns / rows of ONE dense array, so the
// updates of up to LU_WINDOW links can be applied together: a deferred link only updates the strip that becomes the pivot
// block and the panels of the next link (first np_next columns and rows of its trailing matrix), with the panels of all
// links since the last full update, K = kback + np; the link that closes the window updates its whole trailing matrix once
// with that K (up to 128).  C traffic and atomics per flop drop by the window length.
// Kernel: 64 x 64 tile per item and column tile, DMMA as above, K streamed in chunks of 16 through a 3-stage cp.async ring
// (re / im planes; L chunk [k][row] pitch 68, U chunk [column][k] pitch 20: conflict-free copies and fragment reads).
// item.w = column tiles | kind << 8 | cap << 16 (strips: kind 1 / 2, cap = pivot columns of the next link).
// ---------------------------------------------------------------------------------------------
constexpr int SR_KC = 16, SR_LLD = SCHUR_T + 4, SR_ULD = SR_KC + 4;
__host__ __device__ constexpr int sr_stage_doubles(int tn) { return 2 * SR_KC * SR_LLD + 2 * tn * SR_ULD; }
constexpr size_t sr_smem(int tn, int st) { return (size_t)st * sr_stage_doubles(tn) * sizeof(double); }
// MT x NT MMA tiles per warp, WM x WN warps, ST ring stages, TNS = column-tile step of the items.
// 256 threads (4 x 2 warps), 3 stages: (2, 4) full 64 x 64 tile, (1, 4) row strip (32 x 64), (2, 2) column strip (64 x 32).
// 128 threads (4 x 1 warps), 2 stages, 4 CTAs per SM: (2, 4) 64 x 32 tile (full and column strip), (1, 4) row strip (32 x 32).
template <int MT, int NT, int WM, int WN, int ST, int TNS>
__device__ __forceinline__ void schur_ring_body(const LuDev& d, const int4 it, double2* __restrict__ fronts, double* srm) {
    const int s = it.x, i0 = it.y, jfirst = it.z, ntiles = it.w & 0xff;
    constexpr int TM = WM * 8 * MT, TN = WN * 8 * NT;  // rows / columns of the CTA tile
    constexpr int NTH = 32 * WM * WN, SR_ST = ST, SR_STAGE = sr_stage_doubles(TNS);
    const int b = blockIdx.y;
    const int nf = d.ld[s], np = d.np[s], ncb = d.nf[s] - np, kb = d.kback[s];
    const int ktot = kb + np, nk = (ktot + SR_KC - 1) / SR_KC;
    double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dbg = g_schur_dbg;
    const int cap = (it.w >> 16) & 0xff;  // strips: the next link's pivot columns (<= 32)
    const int th = (MT == 1) ? min(cap, ncb - i0) : min(TM, ncb - i0);
    auto tile_w = [&](int j0) { return (((it.w >> 8) & 0xff) == 2) ? min(cap, ncb - j0) : min(TN, ncb - j0); };
    auto load = [&](int i) {
        if (dbg & 8) return;
        const int q = i / nk, kc = i - q * nk;
        const int j0 = jfirst + q * TNS, tw = tile_w(j0);
        double* sLr = srm + (size_t)(i % SR_ST) * SR_STAGE;
        double* sLi = sLr + SR_KC * SR_LLD;
        double* sUr = sLi + SR_KC * SR_LLD;
        double* sUi = sUr + TNS * SR_ULD;
#pragma unroll
        for (int idx = tid; idx < SR_KC * TM; idx += NTH) {  // L21 chunk: TM rows x 16 k, rows fastest (column-major front)
            const int r = idx % TM, kk = idx / TM;
            const int t = kc * SR_KC + kk;
            const bool ok = r < th && t < ktot;
            const double* src = (const double*)(F + (ok ? (int64_t)(np + i0 + r) + (int64_t)(t - kb) * nf : 0));
            cp_async8z(sLr + kk * SR_LLD + r, src, ok);
            cp_async8z(sLi + kk * SR_LLD + r, src + 1, ok);
        }
#pragma unroll
        for (int idx = tid; idx < SR_KC * TN; idx += NTH) {  // U12 chunk: 16 k x TN columns, k fastest
            const int kk = idx % SR_KC, c = idx / SR_KC;
            const int t = kc * SR_KC + kk;
            const bool ok = c < tw && t < ktot;
            const double* src = (const double*)(F + (ok ? (int64_t)(t - kb) + (int64_t)(np + j0 + c) * nf : 0));
            cp_async8z(sUr + c * SR_ULD + kk, src, ok);
            cp_async8z(sUi + c * SR_ULD + kk, src + 1, ok);
        }
    };
    const int total = ntiles * nk;
    for (int i = 0; i < SR_ST - 1; ++i) {
        if (i < total) load(i);
        cp_async_commit();
    }
    const int ar = lane >> 2, ak = lane & 3;
    const int wr = (warp % WM) * 8 * MT, wc = (warp / WM) * 8 * NT;
    double cr[MT][NT][2], ci[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) cr[mt][nt][0] = cr[mt][nt][1] = ci[mt][nt][0] = ci[mt][nt][1] = 0.0;
    int q = 0, kc = 0;
    for (int i = 0; i < total; ++i) {
        cp_async_wait<SR_ST - 2>();
        __syncthreads();  // chunk i has landed for everybody, and everybody is done with the stage that is refilled now
        if (i + SR_ST - 1 < total) load(i + SR_ST - 1);
        cp_async_commit();
        const int j0 = jfirst + q * TNS, tw = tile_w(j0);
        const bool active = wr < th && wc < tw;  // warp-uniform
        if (active && !(dbg & 1)) {
            const double* sLr = srm + (size_t)(i % SR_ST) * SR_STAGE;
            const double* sLi = sLr + SR_KC * SR_LLD;
            const double* sUr = sLi + SR_KC * SR_LLD;
            const double* sUi = sUr + TNS * SR_ULD;
#pragma unroll
            for (int k4 = 0; k4 < SR_KC; k4 += 4) {
                double a_r[MT], a_i[MT], na_i[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    a_r[mt] = sLr[(k4 + ak) * SR_LLD + wr + mt * 8 + ar];
                    a_i[mt] = sLi[(k4 + ak) * SR_LLD + wr + mt * 8 + ar];
                    na_i[mt] = -a_i[mt];
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b_r = sUr[(wc + nt * 8 + ar) * SR_ULD + k4 + ak];
                    const double b_i = sUi[(wc + nt * 8 + ar) * SR_ULD + k4 + ak];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], a_r[mt], b_r);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], a_r[mt], b_i);
                        dmma884(cr[mt][nt][0], cr[mt][nt][1], na_i[mt], b_i);
                        dmma884(ci[mt][nt][0], ci[mt][nt][1], a_i[mt], b_r);
                    }
                }
            }
        }
        if (++kc == nk) {  // tile q complete: C -= acc (L2 reductions, fire and forget), next column tile
            if (active && !(dbg & 2)) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = wc + nt * 8 + 2 * ak + e;
                        if (cc < tw) {
                            double2* col = F + (size_t)(np + i0) + (size_t)(np + j0 + cc) * nf;
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) {
                                const int rr = wr + mt * 8 + ar;
                                if (rr < th) red_sub_c(col + rr, make_double2(cr[mt][nt][e], ci[mt][nt][e]));
                            }
                        }
                    }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) cr[mt][nt][0] = cr[mt][nt][1] = ci[mt][nt][0] = ci[mt][nt][1] = 0.0;
            kc = 0;
            ++q;
        }
    }
}

// item.w = column tiles | kind << 8 | cap << 16: kind 0 full tiles, 1 row strip (32-row tiles), 2 column strip (<= 32 columns)
__global__ void __launch_bounds__(256, 2) lu_schur_ring_kernel(LuDev d, const int4* __restrict__ items, double2* __restrict__ fronts) {
    extern __shared__ double srm[];
    const int4 it = items[blockIdx.x];
    const int kind = (it.w >> 8) & 0xff;
    if (kind == 0) schur_ring_body<2, 4, 4, 2, 3, 64>(d, it, fronts, srm);
    else if (kind == 1) schur_ring_body<1, 4, 4, 2, 3, 64>(d, it, fronts, srm);
    else schur_ring_body<2, 2, 4, 2, 3, 64>(d, it, fronts, srm);
}
// 128 threads, 64 x 32 tiles, two stages: four CTAs per SM, so that the operand loads / epilogues of three CTAs run under the
// MMA phase of the fourth
__global__ void __launch_bounds__(128, 4) lu_schur_ring32_kernel(LuDev d, const int4* __restrict__ items, double2* __restrict__ fronts) {
    extern __shared__ double srm[];
    const int4 it = items[blockIdx.x];
    const int kind = (it.w >> 8) & 0xff;
    if (kind == 1) schur_ring_body<1, 4, 4, 1, 2, 32>(d, it, fronts, srm);
    else schur_ring_body<2, 4, 4, 1, 2, 32>(d, it, fronts, srm);
}

// ---------------------------------------------------------------------------------------------
// solves.  Xp: permuted right-hand sides / solutions, [b][n][k] row-major; W: per-front work rows [b][w_total][k].
// rhs_stride = 0 when all shifts share one right-hand side block (contour integration).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lu_permute_in_kernel(int n, int k, const int* __restrict__ rperm, const double* __restrict__ dr,
                                                            const double2* __restrict__ Bm, size_t rhs_stride, double2* __restrict__ Xp) {
    const int b = blockIdx.y;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * k) return;
    const int i = (int)(idx / k), c = (int)(idx % k);
    const int r = rperm[i];  // operator row that became row i of the factorised matrix
    double2 v = Bm[b * rhs_stride + (size_t)r * k + c];
    if (dr) {
        const double sc = dr[r];
        v.x *= sc;
        v.y *= sc;
    }
    Xp[(size_t)b * n * k + idx] = v;
}

__global__ void __launch_bounds__(256) lu_permute_out_kernel(int n, int k, const int* __restrict__ iperm, const double* __restrict__ dc,
                                                             const double2* __restrict__ Xp, double2* __restrict__ X, size_t x_stride) {
    const int b = blockIdx.y;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * k) return;
    const int i = (int)(idx / k), c = (int)(idx % k);
    double2 v = Xp[(size_t)b * n * k + (size_t)iperm[i] * k + c];
    if (dc) {
        const double sc = dc[i];
        v.x *= sc;
        v.y *= sc;
    }
    X[b * x_stride + idx] = v;
}

constexpr int SOLVE_BIG = 192;    // fronts with more update rows than this get several CTAs in the solves
constexpr int SOLVE_CHUNK = 64;   // update rows per CTA for those
constexpr int SOLVE_TILE = 64;    // update rows staged in shared memory at a time

// shared memory of the solve kernels: sy [np*k] | sT [np*SOLVE_TILE] | sX [SOLVE_TILE*k]
__host__ __device__ inline size_t solve_smem_bytes(int max_np, int k) {
    return ((size_t)max_np * k + (size_t)max_np * SOLVE_TILE + (size_t)SOLVE_TILE * k) * 16;
}

// W[np + r, :] -= L21[r, :] * y1 for the update rows [ra, rb): 64-row tiles of L21 are staged in shared memory with
// coalesced, independent loads (the factors stream from HBM exactly once), then every thread owns one row and CK columns.
template <int CK>
__device__ __forceinline__ void fwd_update_rows(const double2* __restrict__ F, int ld, int np, int ra, int rb, const double2* sy, int k,
                                                double2* __restrict__ Ws, double2* sT) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nck = (k + CK - 1) / CK;
    for (int r0 = ra; r0 < rb; r0 += SOLVE_TILE) {
        const int th = min(SOLVE_TILE, rb - r0);
        __syncthreads();
        for (int idx = tid; idx < np * SOLVE_TILE; idx += nth) {
            const int r = idx % SOLVE_TILE, t = idx / SOLVE_TILE;
            sT[idx] = (r < th) ? F[(size_t)(np + r0 + r) + (size_t)t * ld] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        for (int item = tid; item < SOLVE_TILE * nck; item += nth) {
            const int r = item % SOLVE_TILE, cb = (item / SOLVE_TILE) * CK;
            if (r >= th) continue;
            double2 acc[CK];
#pragma unroll
            for (int c = 0; c < CK; ++c) acc[c] = make_double2(0.0, 0.0);
            for (int t = 0; t < np; ++t) {
                const double2 l = sT[t * SOLVE_TILE + r];
                const double2* yy = sy + t * k + cb;
#pragma unroll
                for (int c = 0; c < CK; ++c)
                    if (cb + c < k) cfma2(acc[c], l, yy[c]);
            }
            double2* w = Ws + (size_t)(np + r0 + r) * k + cb;
#pragma unroll
            for (int c = 0; c < CK; ++c)
                if (cb + c < k) {
                    double2 o = w[c];
                    o.x -= acc[c].x;
                    o.y -= acc[c].y;
                    w[c] = o;
                }
        }
    }
}

// sp[np x k] -= (sign) U12[:, x0:x1] * x2[x0:x1]; tiles of 64 update rows staged in shared memory; the x-range of a
// tile is split over 4 thread groups whose partial sums are added in a fixed order.
template <int CK>
__device__ __forceinline__ void bwd_product(const double2* __restrict__ F, int ld, int np, int x0, int x1, const int* __restrict__ rows,
                                            const double2* __restrict__ Xb, int k, double2* sp, double2* sT, double2* sX, bool subtract) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nck = (k + CK - 1) / CK;
    constexpr int Q = 4;
    for (int xa = x0; xa < x1; xa += SOLVE_TILE) {
        const int tw = min(SOLVE_TILE, x1 - xa);
        __syncthreads();
        for (int idx = tid; idx < np * tw; idx += nth) {
            const int t = idx % np, x = idx / np;
            sT[idx] = F[(size_t)t + (size_t)(np + xa + x) * ld];  // sT[x*np + t]
        }
        for (int idx = tid; idx < tw * k; idx += nth) {
            const int x = idx / k, c = idx % k;
            sX[idx] = Xb[(size_t)rows[xa + x] * k + c];
        }
        __syncthreads();
        const int xq = (tw + Q - 1) / Q;
        const int nitems = np * nck * Q;
        for (int base = 0; base < nitems; base += nth) {  // uniform trip count: barriers inside
            const int item = base + tid;
            const bool ok = item < nitems;
            const int t = item % np, cb = ((item / np) % nck) * CK, q = item / (np * nck);
            double2 acc[CK];
#pragma unroll
            for (int c = 0; c < CK; ++c) acc[c] = make_double2(0.0, 0.0);
            if (ok) {
                const int xe = min(tw, (q + 1) * xq);
                for (int x = q * xq; x < xe; ++x) {
                    const double2 u = sT[x * np + t];
                    const double2* xx = sX + x * k + cb;
#pragma unroll
                    for (int c = 0; c < CK; ++c)
                        if (cb + c < k) cfma2(acc[c], u, xx[c]);
                }
            }
            for (int qq = 0; qq < Q; ++qq) {
                if (ok && q == qq) {
                    double2* y = sp + t * k + cb;
#pragma unroll
                    for (int c = 0; c < CK; ++c)
                        if (cb + c < k) {
                            if (subtract) { y[c].x -= acc[c].x; y[c].y -= acc[c].y; }
                            else { y[c].x += acc[c].x; y[c].y += acc[c].y; }
                        }
                }
                __syncthreads();
            }
        }
    }
}

// forward: one CTA per (front, shift).  CK = right-hand-side columns held in registers per pass.
template <int CK>
__device__ __forceinline__ void forward_front(const LuDev& d, int s, int b, const double2* __restrict__ fronts, const int* __restrict__ piv,
                                              double2* __restrict__ Xp, double2* __restrict__ W, int k, int max_np, double2* sy,
                                              bool update_all) {
    double2* sT = sy + (size_t)max_np * k;
    const int nf = d.nf[s], np = d.np[s], ncb = nf - np, ld = d.ld[s];
    const double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    const int* pv = piv + (size_t)b * d.n + d.sn_ptr[s];
    double2* Wb = W + (size_t)b * d.w_total * k;
    double2* Ws = Wb + (size_t)d.w_off[s] * k;
    double2* Xb = Xp + (size_t)b * d.n * k;
    const int c0 = d.sn_ptr[s];
    const int tid = threadIdx.x;
    // 1. gather: pivot rows from the right-hand side, update rows start at zero; then the children's updates.
    //    An in-place child has already left its update rows in this front's work rows (they are the same memory).
    if (d.has_ip[s]) {
        for (int idx = tid; idx < np * k; idx += blockDim.x) {
            const double2 a = Xb[(size_t)c0 * k + idx], w = Ws[idx];
            sy[idx] = make_double2(a.x + w.x, a.y + w.y);
        }
    } else {
        for (int idx = tid; idx < np * k; idx += blockDim.x) sy[idx] = Xb[(size_t)c0 * k + idx];
        for (int idx = tid; idx < ncb * k; idx += blockDim.x) Ws[(size_t)np * k + idx] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    for (int ci = d.child_ptr[s]; ci < d.child_ptr[s + 1]; ++ci) {
        const int c = d.child_list[ci];
        if (d.in_place[c]) continue;
        const int npc = d.np[c], ncbc = d.nf[c] - npc;
        const int* rel = d.rel + d.rel_ptr[c];
        const double2* Wc = Wb + ((size_t)d.w_off[c] + npc) * k;
        for (int idx = tid; idx < ncbc * k; idx += blockDim.x) {
            const int x = idx / k, col = idx % k;
            const int r = rel[x];
            const double2 v = Wc[idx];
            double2* t = (r < np) ? &sy[r * k + col] : &Ws[(size_t)r * k + col];
            double2 o = *t;
            o.x += v.x;
            o.y += v.y;
            *t = o;
        }
        __syncthreads();
    }
    // 2.+3. y1 = inv(L11) (P b1): one small dense product, no dependency chain (inv(L11) is stored by lu_diag_inv_kernel)
    double2* sTmp = sT + (size_t)max_np * SOLVE_TILE;  // the sX region: np x k permuted right-hand side
    for (int idx = tid; idx < np * k; idx += blockDim.x) sTmp[idx] = sy[pv[idx / k] * k + idx % k];
    for (int idx = tid; idx < np * np; idx += blockDim.x) {
        const int i = idx % np, t = idx / np;
        sT[idx] = (t < i) ? F[(size_t)i + (size_t)t * ld] : make_double2(i == t ? 1.0 : 0.0, 0.0);  // sT[t*np + i] = inv(L11)[i][t]
    }
    __syncthreads();
    for (int idx = tid; idx < np * k; idx += blockDim.x) {
        const int i = idx / k, col = idx % k;
        double2 acc = make_double2(0.0, 0.0);
        for (int t = 0; t <= i; ++t) cfma2(acc, sT[t * np + i], sTmp[t * k + col]);
        sy[idx] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < np * k; idx += blockDim.x) Xb[(size_t)c0 * k + idx] = sy[idx];
    // 4. update rows; big fronts leave this to lu_forward_update_kernel (several CTAs per front) unless the caller owns
    //    the whole chain (lu_forward_chain_kernel)
    if (ncb > SOLVE_BIG && !update_all) return;
    fwd_update_rows<CK>(F, ld, np, 0, ncb, sy, k, Ws, sT);
}

template <int CK>
__global__ void __launch_bounds__(256) lu_forward_kernel(LuDev d, const int* __restrict__ items, const double2* __restrict__ fronts,
                                                         const int* __restrict__ piv, double2* __restrict__ Xp, double2* __restrict__ W, int k,
                                                         int max_np) {
    extern __shared__ double2 sy[];  // np x k pivot rows | L21 tile | staging
    forward_front<CK>(d, items[blockIdx.x], blockIdx.y, fronts, piv, Xp, W, k, max_np, sy, false);
}

// A chain of in-place fronts (the links of a split supernode, e.g. the dense root) handled by ONE CTA per shift: links 2..m
// of the chain are walked inside the kernel, so m-1 levels cost one launch and no inter-CTA synchronisation.
// item = (first index into chain_fronts, number of links)
template <int CK>
__global__ void __launch_bounds__(256) lu_forward_chain_kernel(LuDev d, const int2* __restrict__ items, const int* __restrict__ chain_fronts,
                                                               const double2* __restrict__ fronts, const int* __restrict__ piv,
                                                               double2* __restrict__ Xp, double2* __restrict__ W, int k, int max_np) {
    extern __shared__ double2 sy[];
    const int2 it = items[blockIdx.x];
    for (int q = 0; q < it.y; ++q) {
        forward_front<CK>(d, chain_fronts[it.x + q], blockIdx.y, fronts, piv, Xp, W, k, max_np, sy, true);
        __syncthreads();  // the next link reads the work rows this one just updated
    }
}

// forward, big fronts: item = (front s, first update row r0, end r1); W[np + r, :] -= L21[r, :] * y1 with y1 from Xp
template <int CK>
__global__ void __launch_bounds__(128) lu_forward_update_kernel(LuDev d, const int4* __restrict__ items, const double2* __restrict__ fronts,
                                                                const double2* __restrict__ Xp, double2* __restrict__ W, int k, int max_np) {
    extern __shared__ double2 sy[];
    double2* sT = sy + (size_t)max_np * k;
    const int4 it = items[blockIdx.x];
    const int s = it.x;
    const int b = blockIdx.y;
    const int np = d.np[s], ld = d.ld[s];
    const double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    double2* Ws = W + (size_t)b * d.w_total * k + (size_t)d.w_off[s] * k;
    const double2* Xb = Xp + (size_t)b * d.n * k;
    const int c0 = d.sn_ptr[s];
    for (int idx = threadIdx.x; idx < np * k; idx += blockDim.x) sy[idx] = Xb[(size_t)c0 * k + idx];
    fwd_update_rows<CK>(F, ld, np, it.y, it.z, sy, k, Ws, sT);
}

// backward, big fronts: item = (front s, update rows [x0, x1), slot); part[b][slot][np x k] = U12[:, x0:x1] * x2[x0:x1]
template <int CK>
__global__ void __launch_bounds__(128) lu_backward_partial_kernel(LuDev d, const int4* __restrict__ items, const double2* __restrict__ fronts,
                                                                  const double2* __restrict__ Xp, double2* __restrict__ part, int part_slots,
                                                                  int max_np, int k) {
    extern __shared__ double2 sp[];
    double2* sT = sp + (size_t)max_np * k;
    double2* sX = sT + (size_t)max_np * SOLVE_TILE;
    const int4 it = items[blockIdx.x];
    const int s = it.x;
    const int b = blockIdx.y;
    const int np = d.np[s], ld = d.ld[s];
    const double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    const double2* Xb = Xp + (size_t)b * d.n * k;
    const int* rows = d.rows + d.row_ptr[s] + np;
    for (int idx = threadIdx.x; idx < np * k; idx += blockDim.x) sp[idx] = make_double2(0.0, 0.0);
    bwd_product<CK>(F, ld, np, it.y, it.z, rows, Xb, k, sp, sT, sX, false);
    double2* out = part + ((size_t)b * part_slots + it.w) * (size_t)max_np * k;
    for (int idx = threadIdx.x; idx < np * k; idx += blockDim.x) out[idx] = sp[idx];
}

// backward: x1 = U11^-1 (y1 - U12 x2), one CTA per (front, shift)
template <int CK>
__device__ __forceinline__ void backward_front(const LuDev& d, int s, int b, const double2* __restrict__ fronts, double2* __restrict__ Xp,
                                               const double2* __restrict__ part, int part_slots, int max_np, int k, double2* sy,
                                               bool product_here) {
    double2* sT = sy + (size_t)max_np * k;
    double2* sX = sT + (size_t)max_np * SOLVE_TILE;
    const int nf = d.nf[s], np = d.np[s], ncb = nf - np, ld = d.ld[s];
    const double2* F = fronts + (size_t)b * d.front_total + d.front_off[s];
    double2* Xb = Xp + (size_t)b * d.n * k;
    const int* rows = d.rows + d.row_ptr[s] + np;
    const int c0 = d.sn_ptr[s];
    const int tid = threadIdx.x;
    if (ncb > SOLVE_BIG && !product_here) {
        // update rows [xs, ncb) belong to ancestors above the parent: their products were formed one level earlier by
        // lu_backward_partial_kernel (off the critical path) and are subtracted slot after slot; the rows [0, xs) are the
        // parent's pivot rows, solved by the launch right before this one -- that short product is done here
        const int xs = d.xsplit[s];
        const int nslots = (ncb - xs + SOLVE_CHUNK - 1) / SOLVE_CHUNK;
        const double2* pp = part + ((size_t)b * part_slots + d.bw_slot[s]) * (size_t)max_np * k;
        for (int idx = tid; idx < np * k; idx += blockDim.x) {
            double2 a = Xb[(size_t)c0 * k + idx];
            for (int q = 0; q < nslots; ++q) {
                const double2 v = pp[(size_t)q * max_np * k + idx];
                a.x -= v.x;
                a.y -= v.y;
            }
            sy[idx] = a;
        }
        bwd_product<CK>(F, ld, np, 0, xs, rows, Xb, k, sy, sT, sX, true);
    } else {
        for (int idx = tid; idx < np * k; idx += blockDim.x) sy[idx] = Xb[(size_t)c0 * k + idx];
        bwd_product<CK>(F, ld, np, 0, ncb, rows, Xb, k, sy, sT, sX, true);
    }
    __syncthreads();
    // x1 = inv(U11) * sy
    for (int idx = tid; idx < np * np; idx += blockDim.x) {
        const int i = idx % np, t = idx / np;
        sT[idx] = (t >= i) ? F[(size_t)i + (size_t)t * ld] : make_double2(0.0, 0.0);  // sT[t*np + i] = inv(U11)[i][t]
    }
    for (int idx = tid; idx < np * k; idx += blockDim.x) sX[idx] = sy[idx];
    __syncthreads();
    for (int idx = tid; idx < np * k; idx += blockDim.x) {
        const int i = idx / k, col = idx % k;
        double2 acc = make_double2(0.0, 0.0);
        for (int t = i; t < np; ++t) cfma2(acc, sT[t * np + i], sX[t * k + col]);
        sy[idx] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < np * k; idx += blockDim.x) Xb[(size_t)c0 * k + idx] = sy[idx];
}

template <int CK>
__global__ void __launch_bounds__(256) lu_backward_kernel(LuDev d, const int* __restrict__ items, const double2* __restrict__ fronts,
                                                          double2* __restrict__ Xp, const double2* __restrict__ part, int part_slots,
                                                          int max_np, int k) {
    extern __shared__ double2 sy[];
    backward_front<CK>(d, items[blockIdx.x], blockIdx.y, fronts, Xp, part, part_slots, max_np, k, sy, false);
}

// the same chain, top-down: links m..2 inside one CTA per shift
template <int CK>
__global__ void __launch_bounds__(256) lu_backward_chain_kernel(LuDev d, const int2* __restrict__ items, const int* __restrict__ chain_fronts,
                                                                const double2* __restrict__ fronts, double2* __restrict__ Xp, int max_np, int k) {
    extern __shared__ double2 sy[];
    const int2 it = items[blockIdx.x];
    for (int q = it.y - 1; q >= 0; --q) {
        backward_front<CK>(d, chain_fronts[it.x + q], blockIdx.y, fronts, Xp, nullptr, 0, max_np, k, sy, true);
        __syncthreads();  // the next (lower) link gathers the solution rows written here
    }
}

// ---------------------------------------------------------------------------------------------
// Chains of in-place fronts (the links of a supernode split at 32 pivot columns; the 886-row dense root of gun is 28 of them)
// solved by ONE thread-block cluster per (chain, shift) instead of one launch per link and direction.
// All links of a chain live in one dense column-major array, so the chain is a dense blocked triangular solve.  Link q is
// owned by CTA q mod C of the cluster; the algorithm is right-looking: as soon as the owner has finished the pivot rows of a
// link (one product with the stored inverse of the pivot block) it publishes them in HBM/L2, the cluster synchronises once
// (barrier.cluster, release/acquire), and every CTA applies that link's block column (backward: U, forward: L) to the rows
// of the links it owns -- no reductions across CTAs, one hardware barrier per link, ~5 us per link instead of the
// ~75 us of per-level launches (profiles/r2_contour_latency_steps.txt).  Rows that belong to ancestors above the chain
// (a chain that does not end at the root) are handled in 32-row blocks dealt round-robin to the CTAs.
// Data written by other CTAs of the cluster is read with ld.global.cg (L2), never through the non-coherent path.
// ---------------------------------------------------------------------------------------------
constexpr int CHAIN_C = 8;  // CTAs per cluster (portable maximum)

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// Y[r, :] -= sum_t A(r, t) X[t, :] for r < R <= 32, t < T <= 32.  A(r, t) = A0[r*sr + t*st] (factor entries, read-only);
// X rows: Xg[xrow(t)*k ..] with xrow(t) = xrows ? xrows[t] : x0 + t; Y rows: Yg[(y0 + r)*k ..].  sA, sX: shared staging.
__device__ __forceinline__ void chain_block_update(double2* Yg, int y0, const double2* __restrict__ A0, size_t sr, size_t st, int R, int T,
                                                   const double2* Xg, int x0, const int* __restrict__ xrows, int k, double2* sA, double2* sX) {
    const int tid = threadIdx.x, nth = blockDim.x;
    __syncthreads();
    if (sr == 1) {  // rows contiguous in memory
        for (int idx = tid; idx < R * T; idx += nth) {
            const int r = idx % R, t = idx / R;
            sA[t * 33 + r] = A0[(size_t)r + (size_t)t * st];
        }
    } else {        // columns (t) contiguous in memory
        for (int idx = tid; idx < R * T; idx += nth) {
            const int t = idx % T, r = idx / T;
            sA[t * 33 + r] = A0[(size_t)r * sr + (size_t)t * st];
        }
    }
    for (int idx = tid; idx < T * k; idx += nth) {
        const int t = idx / k, c = idx % k;
        const int xr = xrows ? xrows[t] : x0 + t;
        sX[idx] = __ldcg(Xg + (size_t)xr * k + c);
    }
    __syncthreads();
    for (int idx = tid; idx < R * k; idx += nth) {
        const int r = idx / k, c = idx % k;
        double2 acc = make_double2(0.0, 0.0);
        for (int t = 0; t < T; ++t) cfma2(acc, sA[t * 33 + r], sX[t * k + c]);
        double2* y = Yg + (size_t)(y0 + r) * k + c;
        double2 o = __ldcg(y);
        o.x -= acc.x;
        o.y -= acc.y;
        *y = o;
    }
}

// backward: items[blockIdx.x / C] = (first index into chain_fronts, links); links are stored bottom-up
__global__ void __launch_bounds__(256) lu_backward_chain_cluster_kernel(LuDev d, const int2* __restrict__ items, const int* __restrict__ chain_fronts,
                                                                        const double2* __restrict__ fronts, double2* Xp, int k) {
    extern __shared__ double2 csm[];
    double2* sA = csm;                 // 32 x 33
    double2* sX = sA + 32 * 33;        // 32 x k
    double2* sY = sX + 32 * k;         // 32 x k
    const int2 it = items[blockIdx.x / CHAIN_C];
    const int rank = (int)cluster_cta_rank();
    const int b = blockIdx.y, m = it.y, tid = threadIdx.x, nth = blockDim.x;
    const int* cf = chain_fronts + it.x;
    double2* Xb = Xp + (size_t)b * d.n * k;
    const double2* Fb = fronts + (size_t)b * d.front_total;
    // rows of ancestors above the chain: their solution is final; every owner removes them from its links first
    const int top = cf[m - 1];
    const int nE = d.nf[top] - d.np[top];
    if (nE > 0) {
        for (int i = rank; i < m; i += CHAIN_C) {
            const int s = cf[i];
            const int np = d.np[s], ncb = d.nf[s] - np, ld = d.ld[s];
            const double2* F = Fb + d.front_off[s];
            const int* rows = d.rows + d.row_ptr[s] + np;
            for (int x = ncb - nE; x < ncb; x += 32)
                chain_block_update(Xb, d.sn_ptr[s], F + (size_t)(np + x) * ld, 1, ld, np, min(32, ncb - x), Xb, 0, rows + x, k, sA, sX);
        }
    }
    for (int j = m - 1; j >= 0; --j) {
        const int sj = cf[j];
        const int npj = d.np[sj], c0j = d.sn_ptr[sj];
        if (j % CHAIN_C == rank) {  // x_j = inv(U11) y_j
            const int ld = d.ld[sj];
            const double2* F = Fb + d.front_off[sj];
            __syncthreads();
            for (int idx = tid; idx < npj * npj; idx += nth) {
                const int i = idx % npj, t = idx / npj;
                sA[t * 33 + i] = (t >= i) ? F[(size_t)i + (size_t)t * ld] : make_double2(0.0, 0.0);
            }
            for (int idx = tid; idx < npj * k; idx += nth) sY[idx] = __ldcg(Xb + (size_t)c0j * k + idx);
            __syncthreads();
            for (int idx = tid; idx < npj * k; idx += nth) {
                const int i = idx / k, c = idx % k;
                double2 acc = make_double2(0.0, 0.0);
                for (int t = i; t < npj; ++t) cfma2(acc, sA[t * 33 + i], sY[t * k + c]);
                Xb[(size_t)c0j * k + idx] = acc;
            }
        }
        cluster_sync_all();
        for (int i = rank; i < j; i += CHAIN_C) {  // block column j of U applied to the owned links below it
            const int s = cf[i];
            const int np = d.np[s], ld = d.ld[s], c0 = d.sn_ptr[s];
            const double2* F = Fb + d.front_off[s];
            const int xoff = c0j - (c0 + np);  // link j's pivot columns inside link i's update columns
            chain_block_update(Xb, c0, F + (size_t)(np + xoff) * ld, 1, ld, np, npj, Xb, c0j, nullptr, k, sA, sX);
        }
    }
}

// forward: pivot rows of a link = right-hand side + what the links below left in its work rows; update rows live in W
__global__ void __launch_bounds__(256) lu_forward_chain_cluster_kernel(LuDev d, const int2* __restrict__ items, const int* __restrict__ chain_fronts,
                                                                       const double2* __restrict__ fronts, const int* __restrict__ piv, double2* Xp,
                                                                       double2* W, int k) {
    extern __shared__ double2 csm[];
    double2* sA = csm;
    double2* sX = sA + 32 * 33;
    double2* sY = sX + 32 * k;
    const int2 it = items[blockIdx.x / CHAIN_C];
    const int rank = (int)cluster_cta_rank();
    const int b = blockIdx.y, m = it.y, tid = threadIdx.x, nth = blockDim.x;
    const int* cf = chain_fronts + it.x;
    double2* Xb = Xp + (size_t)b * d.n * k;
    double2* Wb = W + (size_t)b * d.w_total * k;
    const double2* Fb = fronts + (size_t)b * d.front_total;
    const int top = cf[m - 1];
    const int nE = d.nf[top] - d.np[top];
    for (int q = 0; q < m; ++q) {
        const int s = cf[q];
        const int np = d.np[s], ncb = d.nf[s] - np, ld = d.ld[s], c0 = d.sn_ptr[s];
        const double2* F = Fb + d.front_off[s];
        if (q % CHAIN_C == rank) {  // y1 = inv(L11) P (b1 + work rows)
            const int* pv = piv + (size_t)b * d.n + c0;
            const double2* Ws = Wb + (size_t)d.w_off[s] * k;
            __syncthreads();
            for (int idx = tid; idx < np * k; idx += nth) {
                const double2 a = __ldcg(Xb + (size_t)c0 * k + idx), w = __ldcg(Ws + idx);
                sY[idx] = make_double2(a.x + w.x, a.y + w.y);
            }
            for (int idx = tid; idx < np * np; idx += nth) {
                const int i = idx % np, t = idx / np;
                sA[t * 33 + i] = (t < i) ? F[(size_t)i + (size_t)t * ld] : make_double2(i == t ? 1.0 : 0.0, 0.0);
            }
            __syncthreads();
            for (int idx = tid; idx < np * k; idx += nth) sX[idx] = sY[pv[idx / k] * k + idx % k];
            __syncthreads();
            for (int idx = tid; idx < np * k; idx += nth) {
                const int i = idx / k, c = idx % k;
                double2 acc = make_double2(0.0, 0.0);
                for (int t = 0; t <= i; ++t) cfma2(acc, sA[t * 33 + i], sX[t * k + c]);
                Xb[(size_t)c0 * k + idx] = acc;
            }
        }
        cluster_sync_all();
        // block column q of L applied to the rows below: the pivot rows of the later links (owner = link mod C) ...
        const size_t w0 = (size_t)d.w_off[s] + np;  // first update row of this link in the work rows
        for (int i = q + 1; i < m; ++i) {
            if (i % CHAIN_C != rank) continue;
            const int si = cf[i];
            const int xoff = d.sn_ptr[si] - (c0 + np);
            chain_block_update(Wb + w0 * k, xoff, F + (size_t)(np + xoff), 1, ld, d.np[si], np, Xb, c0, nullptr, k, sA, sX);
        }
        // ... and the rows of the ancestors above the chain, 32 at a time, dealt round-robin
        for (int x = ncb - nE, blk = 0; x < ncb; x += 32, ++blk) {
            if ((blk + q) % CHAIN_C != rank) continue;
            chain_block_update(Wb + w0 * k, x, F + (size_t)(np + x), 1, ld, min(32, ncb - x), np, Xb, c0, nullptr, k, sA, sX);
        }
    }
}

// S[j] += sum_b wgt[b*mg + j] * X[b]   (contour moments, method_contour_common.jl:86-90), elementwise n*k
__global__ void contour_accumulate_kernel(size_t nk, int nb, int mg, const double2* __restrict__ X, size_t x_stride,
                                                                 const double2* __restrict__ wgt, double2* __restrict__ S) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nk) return;
    for (int j = 0; j < mg; ++j) {
        double2 acc = S[(size_t)j * nk + idx];
        for (int b = 0; b < nb; ++b) cfma2(acc, wgt[b * mg + j], X[b * x_stride + idx]);
        S[(size_t)j * nk + idx] = acc;
    }
}

// gather / scatter of column windows between a wide row-major block and the contiguous solve buffers
__global__ void __launch_bounds__(256) cols_gather_kernel(int64_t n, int nc, const double2* __restrict__ src, int lds, double2* __restrict__ dst) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * nc) return;
    dst[idx] = src[(size_t)(idx / nc) * lds + idx % nc];
}
__global__ void __launch_bounds__(256) cols_scatter_kernel(int64_t n, int nc, const double2* __restrict__ src, double2* __restrict__ dst, int ldd,
                                                           double2 alpha) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * nc) return;
    const double2 v = src[idx];
    dst[(size_t)(idx / nc) * ldd + idx % nc] = make_double2(alpha.x * v.x - alpha.y * v.y, alpha.x * v.y + alpha.y * v.x);
}

// R = B - R  (residual from M*X), elementwise
__global__ void __launch_bounds__(256) residual_kernel(size_t count, const double2* __restrict__ Bm, double2* __restrict__ R) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const double2 bv = Bm[idx], r = R[idx];
    R[idx] = make_double2(bv.x - r.x, bv.y - r.y);
}
__global__ void __launch_bounds__(256) axpy_kernel(size_t count, const double2* __restrict__ D, double2* __restrict__ X) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    X[idx].x += D[idx].x;
    X[idx].y += D[idx].y;
}
// per-column max |.| of a row-major n x k block -> out[k] (as ordered uint64 bits)
__global__ void __launch_bounds__(256) colmax_kernel(int n, int k, const double2* __restrict__ A, unsigned long long* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * k) return;
    const double a = cabs1(A[idx]);
    atomicMax(out + idx % k, (unsigned long long)__double_as_longlong(a));
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <class T>
static cudaError_t upload(DevBuf<T>& buf, const std::vector<T>& v) {
    cudaError_t e = buf.alloc(std::max<size_t>(v.size(), 1));
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpy(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

static LuOptions g_default_opt;

// rowmap / dr / dc: optional static-pivoting data from max_product_matching (row i of the operator -> row rowmap[i])
static int lu_symbolic_build(const nepb_spmf* h, const int32_t* rowmap, const double* dr, const double* dc, LuSymbolicDev** out) {
    nepb_spmf* hm = const_cast<nepb_spmf*>(h);
    NEPB_CHECK_ARG(h->n < (int64_t)1 << 31, "n too large");
    LuSymbolicDev* sd = new LuSymbolicDev();
    LuOptions opt = hm->lu_opt_set ? hm->lu_opt : g_default_opt;
    if (const char* e = getenv("NEPB_LU_MAXNP")) opt.max_np = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("NEPB_LU_RELAX")) opt.relax_leaf = std::max(1, atoi(e));
    if (const char* e = getenv("NEPB_LU_ORDERING")) opt.ordering = atoi(e);
    opt.max_np = std::max(1, std::min(32, opt.max_np));  // the pivot-block kernels hold a 32 x 32 block in one warp's lanes
    int rc = lu_symbolic_analyse((int)h->n, h->h_rowptr, h->h_colind, hm->lu_user_perm.empty() || rowmap ? nullptr : hm->lu_user_perm.data(), opt, sd->S, rowmap);
    if (rc) {
        delete sd;
        return rc;
    }
    const LuSymbolic& S = sd->S;
    const int ns = S.nsuper;
    std::vector<int32_t> nf(ns), np(ns), child_ptr(ns + 1, 0), child_list;
    for (int s = 0; s < ns; ++s) {
        nf[s] = (int)(S.row_ptr[s + 1] - S.row_ptr[s]);
        np[s] = S.sn_ptr[s + 1] - S.sn_ptr[s];
        if (S.sn_parent[s] >= 0) child_ptr[S.sn_parent[s] + 1]++;
    }
    for (int s = 0; s < ns; ++s) child_ptr[s + 1] += child_ptr[s];
    child_list.resize(child_ptr[ns]);
    {
        std::vector<int32_t> cur(child_ptr.begin(), child_ptr.end() - 1);
        for (int s = 0; s < ns; ++s)
            if (S.sn_parent[s] >= 0) child_list[cur[S.sn_parent[s]]++] = s;  // ascending child index: fixed order
    }
    // per-level work lists
    sd->lv.resize(S.nlevels);
    std::vector<int32_t> fr_items;
    std::vector<int4> ea_items, pn_items, sc_items, sp_items, fu_items, bp_items;
    std::vector<EaRec> ea_recs;
    std::vector<int32_t> bw_slot(ns, 0), xsplit(ns, 0);
    sd->part_slots = 0;
    for (int s = 0; s < ns; ++s) {
        // update rows of s that are pivot columns of its parent come first (rel is increasing): [0, xsplit)
        const int par = S.sn_parent[s];
        const int ncb = nf[s] - np[s];
        if (par < 0 || ncb == 0) continue;
        const int64_t r0 = S.rel_ptr[s], r1 = S.rel_ptr[s + 1];
        if (r1 - r0 == ncb) {
            const int32_t* rel = S.rel.data() + r0;
            xsplit[s] = (int)(std::lower_bound(rel, rel + ncb, np[par]) - rel);
        } else {
            xsplit[s] = std::min(ncb, (int)np[par]);
        }
    }
    // chains of in-place fronts: the tail links (2..m) are walked by one CTA per shift in the solves
    std::vector<char> in_tail(ns, 0);
    std::vector<int32_t> chain_fronts, sfr_items, sfr_tail;
    std::vector<int4> fu_tail;
    std::vector<std::vector<int2>> fc_of_level(S.nlevels), bc_of_level(S.nlevels);
    // measured on gun (profiles/): a single CTA per shift cannot stream a big chain's factors fast enough (16.0 -> 25.3 ms for
    // 16 nodes), so the chain walk is opt-in (NEPB_LU_CHAINS=1) until it is double-buffered / cluster-wide
    // round 2: the chain tails are solved by one thread-block cluster per (chain, shift) (lu_*_chain_cluster_kernel); chains with
    // fewer than 8 tail links keep the per-level kernels.  NEPB_LU_CHAINS=0 switches the chain kernels off.
    static const bool use_chains = !(getenv("NEPB_LU_CHAINS") && atoi(getenv("NEPB_LU_CHAINS")) == 0);
    for (int s = 0; s < ns && use_chains; ++s) {
        if (!S.in_place_child[s] || S.has_in_place_child[s]) continue;  // not the first link of a chain
        int cnt = 0;
        for (int cur = s; S.in_place_child[cur]; cur = S.sn_parent[cur]) ++cnt;
        // measured on gun (profiles/r2_chain_solves.txt): a cluster launch costs ~10 us per link plus the rows above the chain, and a
        // short chain inside a level that has other fronts only adds a launch to that level
        static const int min_links = getenv("NEPB_LU_CHAIN_MIN") ? atoi(getenv("NEPB_LU_CHAIN_MIN")) : 8;
        if (cnt < min_links) continue;
        const int first = (int)chain_fronts.size();
        int cur = s;
        while (S.in_place_child[cur]) {
            cur = S.sn_parent[cur];
            chain_fronts.push_back(cur);
            in_tail[cur] = 1;
        }
        // both directions are launched at the level of the LAST link: the forward chain needs the panels of all its links (the
        // pipelined factor + forward sequence enqueues level l of the forward solve once the panels of level l are final)
        fc_of_level[S.level[chain_fronts[first + cnt - 1]]].push_back(make_int2(first, cnt));
        bc_of_level[S.level[chain_fronts[first + cnt - 1]]].push_back(make_int2(first, cnt));
        if (getenv("NEPB_LU_DEBUG")) {
            const int top = chain_fronts[first + cnt - 1];
            fprintf(stderr, "[lu] chain: first link front %d (nf %d, level %d), %d tail links up to level %d, %d rows above the chain\n", s, nf[s],
                    S.level[s], cnt, S.level[top], nf[top] - np[top]);
        }
    }
    // delayed Schur updates (lu_schur_ring_kernel): windows of up to LU_WINDOW consecutive links of a chain of in-place fronts
    std::vector<int32_t> kback(ns, 0);
    std::vector<char> defer(ns, 0);
    static const bool use_ring = !(getenv("NEPB_LU_SCHUR_RING") && atoi(getenv("NEPB_LU_SCHUR_RING")) == 0) && !getenv("NEPB_LU_SCHUR_SIMPLE");
    static const int window = std::max(1, std::min(7, getenv("NEPB_LU_WINDOW") ? atoi(getenv("NEPB_LU_WINDOW")) : 4));  // kback + np <= 224: the strip caps are 8-bit
    sd->schur_ring = use_ring && S.max_np <= 32;
    static const int ring_tn = (getenv("NEPB_LU_SCHUR_TN") && atoi(getenv("NEPB_LU_SCHUR_TN")) == 64) ? 64 : 32;
    sd->schur_tn = sd->schur_ring ? ring_tn : SCHUR_T;
    if (sd->schur_ring && window > 1)
        for (int s = 0; s < ns; ++s) {
            if (!S.in_place_child[s] || S.has_in_place_child[s]) continue;  // not the first link of a chain
            int pos = 0, kb = 0;
            for (int cur = s;; cur = S.sn_parent[cur]) {
                kback[cur] = kb;
                const bool has_next = S.in_place_child[cur];
                if (has_next && pos < window - 1) {
                    defer[cur] = 1;
                    kb += np[cur];
                    ++pos;
                } else {
                    kb = 0;
                    pos = 0;
                }
                if (!has_next) break;
            }
        }
    std::vector<int2> fc_items, bc_items;
    int max_level_slots = 0;
    for (int l = 0; l < S.nlevels; ++l) {
        auto& L = sd->lv[l];
        L.front_begin = (int)fr_items.size();
        L.ea_begin = (int)ea_items.size();
        L.pn_begin = (int)pn_items.size();
        L.sc_begin = (int)sc_items.size();
        L.sp_begin = (int)sp_items.size();
        L.fu_begin = (int)fu_items.size();
        L.bp_begin = (int)bp_items.size();
        L.sfr_begin = (int)sfr_items.size();
        L.fc_begin = (int)fc_items.size();
        L.bc_begin = (int)bc_items.size();
        for (auto& x : fc_of_level[l]) fc_items.push_back(x);
        for (auto& x : bc_of_level[l]) bc_items.push_back(x);
        L.fc_count = (int)fc_items.size() - L.fc_begin;
        L.bc_count = (int)bc_items.size() - L.bc_begin;
        int slots = 0;
        std::vector<int32_t> fl(S.level_list.begin() + S.level_ptr[l], S.level_list.begin() + S.level_ptr[l + 1]);
        std::stable_sort(fl.begin(), fl.end(), [&](int a, int b) { return nf[a] > nf[b]; });  // big fronts first
        for (int s : fl) {
            fr_items.push_back(s);
            const int ncb = nf[s] - np[s];
            bool needs_ea = false;
            for (int ci = child_ptr[s]; ci < child_ptr[s + 1]; ++ci)
                if (!S.in_place_child[child_list[ci]] && nf[child_list[ci]] > np[child_list[ci]]) needs_ea = true;
            if (needs_ea) {
                // slab width: keep roughly <= 4k entries of child data per CTA (short CTAs: the kernel sits on the critical path)
                int slab = nf[s];
                if (nf[s] > 64) slab = std::max(8, (int)(4096 / nf[s]));
                for (int j0 = 0; j0 < nf[s]; j0 += slab) {
                    const int j1 = std::min(nf[s], j0 + slab);
                    const int rec0 = (int)ea_recs.size();
                    for (int ci = child_ptr[s]; ci < child_ptr[s + 1]; ++ci) {
                        const int c = child_list[ci];
                        const int ncbc = nf[c] - np[c];
                        if (ncbc == 0 || S.in_place_child[c]) continue;  // in-place child: its Schur update already landed in this front
                        const int32_t* rel = S.rel.data() + S.rel_ptr[c];
                        const int xa = (int)(std::lower_bound(rel, rel + ncbc, j0) - rel);
                        const int xb = (int)(std::lower_bound(rel, rel + ncbc, j1) - rel);
                        if (xb <= xa) continue;
                        EaRec r;
                        r.child_off = S.front_off[c];
                        r.rel_off = S.rel_ptr[c];
                        r.ldc = S.front_ld[c];
                        r.npc = np[c];
                        r.ncb = ncbc;
                        r.xa = xa;
                        r.xb = xb;
                        r.pad = 0;
                        ea_recs.push_back(r);
                    }
                    const int cnt = (int)ea_recs.size() - rec0;
                    if (cnt) ea_items.push_back(make_int4(s, j0, rec0, cnt));
                }
            }
            for (int t0 = 0; t0 < ncb; t0 += PANEL_T) {
                pn_items.push_back(make_int4(s, 0, t0, 0));
                pn_items.push_back(make_int4(s, 1, t0, 0));
            }
            const int tn = sd->schur_tn;  // column-tile step of the Schur items
            if (defer[s]) {
                // deferred link: only the strip that becomes the next link's pivot block and panels (its first npn columns, all
                // rows; its first npn rows, the remaining columns)
                const int npn = np[S.sn_parent[s]];  // <= 32 = the strip tiles' short side
                for (int i0 = 0; i0 < ncb; i0 += SCHUR_T) {  // column strip: 64 x 32 tiles
                    sc_items.push_back(make_int4(s, i0, 0, 1 | (2 << 8) | (npn << 16)));
                    sp_items.push_back(make_int4(s, i0, 0, 1 | (2 << 8) | (npn << 16)));
                }
                // row strip: 32 x 64 tiles over the columns >= npn (the corner belongs to the column strip)
                const int grp = SCHUR_GROUP * (SCHUR_T / tn);  // column tiles per grouped item: the same 128 columns for both widths
                for (int j0 = npn; j0 < ncb; j0 += tn) sc_items.push_back(make_int4(s, 0, j0, 1 | (1 << 8) | (npn << 16)));
                for (int j0 = npn; j0 < ncb; j0 += tn * grp)
                    sp_items.push_back(make_int4(s, 0, j0, std::min(grp, (ncb - j0 + tn - 1) / tn) | (1 << 8) | (npn << 16)));
            } else {
                const int grp = SCHUR_GROUP * (SCHUR_T / tn);
                for (int j0 = 0; j0 < ncb; j0 += tn)
                    for (int i0 = 0; i0 < ncb; i0 += SCHUR_T) sc_items.push_back(make_int4(s, i0, j0, 1));
                for (int j0 = 0; j0 < ncb; j0 += tn * grp)
                    for (int i0 = 0; i0 < ncb; i0 += SCHUR_T)
                        sp_items.push_back(make_int4(s, i0, j0, std::min(grp, (ncb - j0 + tn - 1) / tn)));
            }
            (in_tail[s] ? sfr_tail : sfr_items).push_back(s);
            if (ncb > SOLVE_BIG)
                for (int r0 = 0; r0 < ncb; r0 += SOLVE_CHUNK)
                    (in_tail[s] ? fu_tail : fu_items).push_back(make_int4(s, r0, std::min(ncb, r0 + SOLVE_CHUNK), 0));
            if (ncb > SOLVE_BIG && !in_tail[s]) {
                // backward partial products over the rows of the ancestors above the parent; levels alternate between two
                // slot buffers because these items run while the level above still reads its own slots
                bw_slot[s] = slots;
                for (int r0 = xsplit[s]; r0 < ncb; r0 += SOLVE_CHUNK) bp_items.push_back(make_int4(s, r0, std::min(ncb, r0 + SOLVE_CHUNK), slots++));
            }
        }
        max_level_slots = std::max(max_level_slots, slots);
        // chain tails go behind the other fronts of the level: the per-level forward kernels of the pipelined factor + solve
        // sequence take all of them, everything else stops before the tails (the chain kernels own those)
        L.fu_count = (int)fu_items.size() - L.fu_begin;
        L.sfr_count = (int)sfr_items.size() - L.sfr_begin;
        fu_items.insert(fu_items.end(), fu_tail.begin(), fu_tail.end());
        sfr_items.insert(sfr_items.end(), sfr_tail.begin(), sfr_tail.end());
        L.fu_count_all = (int)fu_items.size() - L.fu_begin;
        L.sfr_count_all = (int)sfr_items.size() - L.sfr_begin;
        fu_tail.clear();
        sfr_tail.clear();
        L.bp_count = (int)bp_items.size() - L.bp_begin;
        L.max_np = 1;
        for (int s2 : fl) L.max_np = std::max(L.max_np, (int)np[s2]);
        L.front_count = (int)fr_items.size() - L.front_begin;
        L.ea_count = (int)ea_items.size() - L.ea_begin;
        L.pn_count = (int)pn_items.size() - L.pn_begin;
        L.sc_count = (int)sc_items.size() - L.sc_begin;
        L.sp_count = (int)sp_items.size() - L.sp_begin;
    }
    cudaError_t e = cudaSuccess;
#define UP(buf, vec) if (e == cudaSuccess) e = upload(buf, vec)
    UP(sd->front_off, S.front_off);
    UP(sd->nf, nf);
    UP(sd->np, np);
    UP(sd->ld, S.front_ld);
    UP(sd->in_place, S.in_place_child);
    UP(sd->row_ptr, S.row_ptr);
    UP(sd->rows, S.rows);
    UP(sd->rel_ptr, S.rel_ptr);
    UP(sd->rel, S.rel);
    UP(sd->sn_ptr, S.sn_ptr);
    UP(sd->w_off, S.w_off);
    UP(sd->child_ptr, child_ptr);
    UP(sd->child_list, child_list);
    UP(sd->a_pos, S.a_pos);
    UP(sd->perm, S.perm);
    UP(sd->iperm, S.iperm);
    {
        // right-hand side gather: row i of the factorised matrix is operator row rperm[i]
        std::vector<int32_t> rperm(S.perm), inv;
        if (rowmap) {
            inv.resize(S.n);
            for (int i = 0; i < S.n; ++i) inv[rowmap[i]] = i;
            for (int i = 0; i < S.n; ++i) rperm[i] = inv[S.perm[i]];
        }
        UP(sd->rperm, rperm);
        if (dr && dc) {
            std::vector<double> vr(dr, dr + S.n), vc(dc, dc + S.n), sc(S.nnz);
            for (int i = 0; i < S.n; ++i)
                for (int e = h->h_rowptr[i]; e < h->h_rowptr[i + 1]; ++e) sc[e] = dr[i] * dc[h->h_colind[e]];
            UP(sd->dr, vr);
            UP(sd->dc, vc);
            UP(sd->a_scale, sc);
        }
    }
    UP(sd->fr_items, fr_items);
    UP(sd->ea_items, ea_items);
    UP(sd->ea_recs, ea_recs);
    UP(sd->pn_items, pn_items);
    UP(sd->sc_items, sc_items);
    UP(sd->sp_items, sp_items);
    UP(sd->fu_items, fu_items);
    // odd levels use the second half of the slot buffer
    sd->part_slots = 2 * std::max(max_level_slots, 1);
    for (int l = 1; l < S.nlevels; l += 2) {
        const auto& L = sd->lv[l];
        for (int i = L.bp_begin; i < L.bp_begin + L.bp_count; ++i) bp_items[i].w += sd->part_slots / 2;
        for (int i = L.sfr_begin; i < L.sfr_begin + L.sfr_count; ++i) {
            const int fs = sfr_items[i];
            if (nf[fs] - np[fs] > SOLVE_BIG) bw_slot[fs] += sd->part_slots / 2;
        }
    }
    UP(sd->bp_items, bp_items);
    UP(sd->bw_slot, bw_slot);
    UP(sd->xsplit, xsplit);
    UP(sd->sfr_items, sfr_items);
    UP(sd->chain_fronts, chain_fronts);
    UP(sd->fc_items, fc_items);
    UP(sd->bc_items, bc_items);
    UP(sd->has_ip, S.has_in_place_child);
    UP(sd->kback, kback);
#undef UP
    if (e != cudaSuccess) {
        set_error("CUDA error while uploading the LU symbolic data: %s", cudaGetErrorString(e));
        delete sd;
        return e == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;
    }
    LuDev& d = sd->dev;
    d.front_off = sd->front_off.p;
    d.nf = sd->nf.p;
    d.np = sd->np.p;
    d.ld = sd->ld.p;
    d.in_place = sd->in_place.p;
    d.has_ip = sd->has_ip.p;
    d.kback = sd->kback.p;
    d.bw_slot = sd->bw_slot.p;
    d.xsplit = sd->xsplit.p;
    d.row_ptr = sd->row_ptr.p;
    d.rows = sd->rows.p;
    d.rel_ptr = sd->rel_ptr.p;
    d.rel = sd->rel.p;
    d.sn_ptr = sd->sn_ptr.p;
    d.w_off = sd->w_off.p;
    d.child_ptr = sd->child_ptr.p;
    d.child_list = sd->child_list.p;
    d.front_total = S.front_total;
    d.w_total = S.w_total;
    d.n = S.n;
    // opt-in shared memory sizes
    const int mnp = S.max_np;
    sd->smem_diag = (size_t)mnp * (mnp + 1) * 16;
    sd->smem_panel = ((size_t)mnp * (mnp + 1) + (size_t)PANEL_T * (mnp + 1)) * 16;
    sd->smem_schur = (size_t)mnp * (2 * SCHUR_T + 1) * 16;
    sd->smem_schur_pipe = (size_t)mnp * (3 * SCHUR_T + 2) * 16;
    sd->schur_pipe = sd->smem_schur_pipe <= 110 * 1024 && !getenv("NEPB_LU_SCHUR_SIMPLE");
    {   // DMMA Schur update (default; NEPB_LU_SCHUR_DMMA=0 selects the DFMA kernels)
        const int mpad = (mnp + 3) & ~3;
        sd->smem_schur_dmma = ((size_t)2 * mpad * SD_LLD + (size_t)2 * SD_UBUF) * sizeof(double);
        const char* e = getenv("NEPB_LU_SCHUR_DMMA");
        sd->schur_dmma = mnp <= 32 && !(e && atoi(e) == 0) && !getenv("NEPB_LU_SCHUR_SIMPLE");
        cudaFuncSetAttribute(lu_schur_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sr_smem(64, 3) + 64 * 1024);
        cudaFuncSetAttribute(lu_schur_ring32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sr_smem(32, 2) + 64 * 1024);
        if (getenv("NEPB_LU_SCHUR_DBG")) {
            const int v = atoi(getenv("NEPB_LU_SCHUR_DBG"));
            cudaMemcpyToSymbol(g_schur_dbg, &v, sizeof(int));
        }
        if (sd->schur_dmma) {
            cudaFuncSetAttribute(lu_schur_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sd->smem_schur_dmma);
            cudaFuncSetAttribute(lu_schur_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sd->smem_schur_dmma);
        }
    }
    cudaFuncSetAttribute(lu_diag_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
    cudaFuncSetAttribute(lu_schur_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sd->smem_schur_pipe);
    cudaFuncSetAttribute(lu_panel_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM);
    cudaFuncSetAttribute(lu_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sd->smem_schur);
#define NEPB_SOLVE_ATTR(CK_)                                                                                    \
    cudaFuncSetAttribute(lu_forward_kernel<CK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);       \
    cudaFuncSetAttribute(lu_backward_kernel<CK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);       \
    cudaFuncSetAttribute(lu_forward_update_kernel<CK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    cudaFuncSetAttribute(lu_forward_chain_kernel<CK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  \
    cudaFuncSetAttribute(lu_backward_chain_kernel<CK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    cudaFuncSetAttribute(lu_backward_partial_kernel<CK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    NEPB_SOLVE_ATTR(1) NEPB_SOLVE_ATTR(4) NEPB_SOLVE_ATTR(8) NEPB_SOLVE_ATTR(10) NEPB_SOLVE_ATTR(16)
    cudaFuncSetAttribute(lu_forward_chain_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(lu_backward_chain_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
#undef NEPB_SOLVE_ATTR
    *out = sd;
    return NEPB_OK;
}

static std::mutex g_sym_mtx;

int lu_symbolic_get(const nepb_spmf* h, LuSymbolicDev** out) {
    std::lock_guard<std::mutex> lock(g_sym_mtx);
    nepb_spmf* hm = const_cast<nepb_spmf*>(h);
    if (hm->lu_prefer_matched && !hm->lu_matched.empty()) {
        *out = (LuSymbolicDev*)hm->lu_matched.back();
        return NEPB_OK;
    }
    if (!hm->lu_symbolic) {
        LuSymbolicDev* sd = nullptr;
        int rc = lu_symbolic_build(h, nullptr, nullptr, nullptr, &sd);
        if (rc) return rc;
        hm->lu_symbolic = sd;
    }
    *out = (LuSymbolicDev*)hm->lu_symbolic;
    return NEPB_OK;
}

void lu_symbolic_release(void* p) { delete (LuSymbolicDev*)p; }

__global__ void __launch_bounds__(256) lu_absval_kernel(int64_t nnz, int p, int ca, const double* __restrict__ vals,
                                                        const double2* __restrict__ coef, double* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int vw = ca ? 2 * p : p;
    const double* v = vals + (size_t)e * vw;
    double2 m = make_double2(0.0, 0.0);
    for (int i = 0; i < p; ++i) {
        const double2 x = ca ? make_double2(v[2 * i], v[2 * i + 1]) : make_double2(v[i], 0.0);
        cfma2(m, coef[i], x);
    }
    out[e] = hypot(m.x, m.y);
}

// Static pivoting for an operator whose diagonal does not carry the weight at this shift: maximum-product matching and
// I-matrix scaling of M(sigma) = sum_i coef_i A_i (lu_matching.cpp), then a fresh symbolic analysis of the row-permuted
// pattern.  The result becomes the operator's preferred analysis (later factorisations start from it); earlier analyses
// stay alive because existing factorisations point to them.
int lu_symbolic_make_matched(const nepb_spmf* h, const double* coef, LuSymbolicDev** out) {
    nepb_spmf* hm = const_cast<nepb_spmf*>(h);
    DevBuf<double> d_abs, d_coef;
    NEPB_CUDA(d_abs.alloc((size_t)h->nnz));
    NEPB_CUDA(d_coef.alloc((size_t)2 * h->p));
    NEPB_CUDA(cudaMemcpyAsync(d_coef.p, coef, sizeof(double) * 2 * h->p, cudaMemcpyHostToDevice, stream()));
    NEPB_LAUNCH(lu_absval_kernel, (unsigned)((h->nnz + 255) / 256), 256, 0, h->nnz, h->p, h->is_complex, h->d_vals.p, (const double2*)d_coef.p, d_abs.p);
    NEPB_LAUNCH_CHECK();
    std::vector<double> absval((size_t)h->nnz), dr, dc;
    NEPB_CUDA(cudaMemcpyAsync(absval.data(), d_abs.p, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, stream()));
    NEPB_CUDA(cudaStreamSynchronize(stream()));
    std::vector<int32_t> rowmap;
    const int matched = max_product_matching((int)h->n, h->h_rowptr, h->h_colind, absval.data(), rowmap, dr, dc);
    if (matched < (int)h->n) {
        set_error("M(sigma) is structurally singular: only %d of %lld rows can be matched to a nonzero", matched, (long long)h->n);
        return NEPB_E_SINGULAR;
    }
    LuSymbolicDev* sd = nullptr;
    int rc = lu_symbolic_build(h, rowmap.data(), dr.data(), dc.data(), &sd);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_sym_mtx);
    hm->lu_matched.push_back(sd);
    hm->lu_prefer_matched = true;
    *out = sd;
    return NEPB_OK;
}

__global__ void lu_info_init_kernel(int nb, LuInfo* __restrict__ info) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    info[b].amax_bits = 0;
    info[b].amax_plain_bits = 0;
    info[b].minpiv_bits = 0x7ff0000000000000ULL;  // +inf
    info[b].flags = 0;
    info[b].nperturbed = 0;
}

// numeric factorisation of lu->nb shifts into lu->fronts (coefficients already on the device); device work only, so the
// sequence can be captured into a CUDA graph
// Timing experiments only (results are wrong when a bit is set): NEPB_LU_SKIP bit0 Schur update, bit1 extend-add, bit2 panels,
// bit3 forward solve, bit4 backward solve, bit5 diag, bit6 zero-fill of the fronts.
static int lu_skip_mask() {
    static const int m = getenv("NEPB_LU_SKIP") ? atoi(getenv("NEPB_LU_SKIP")) : 0;
    return m;
}

static int factor_prologue(nepb_lu* lu) {
    const nepb_spmf* h = lu->op;
    LuSymbolicDev* sd = lu->sym;
    const LuSymbolic& S = sd->S;
    const int nb = lu->nb;
    if (!(lu_skip_mask() & 64)) NEPB_CUDA(cudaMemsetAsync(lu->fronts.p, 0, sizeof(double) * 2 * (size_t)nb * S.front_total, stream()));
    NEPB_LAUNCH(lu_info_init_kernel, (nb + 127) / 128, 128, 0, nb, lu->info.p);
    dim3 grid((unsigned)((h->nnz + 255) / 256), nb);
    NEPB_LAUNCH(lu_assemble_kernel, grid, 256, 0, h->nnz, h->p, h->is_complex, sd->a_pos.p, sd->a_scale.p, h->d_vals.p, (const double2*)lu->coef.p,
                (double2*)lu->fronts.p, S.front_total, lu->info.p);
    return NEPB_OK;
}

// level l up to the panels: after this the factors L, U of the level's fronts are final (the Schur update only touches
// the contribution blocks)
static void factor_level_panels(nepb_lu* lu, int l) {
    LuSymbolicDev* sd = lu->sym;
    const int nb = lu->nb;
    double2* F = (double2*)lu->fronts.p;
    const auto& L = sd->lv[l];
    const int skip = lu_skip_mask();
    if (L.ea_count && !(skip & 2)) NEPB_LAUNCH(lu_extend_add_kernel, dim3(L.ea_count, nb), 256, 0, sd->dev, sd->ea_items.p + L.ea_begin, sd->ea_recs.p, F);
    if (!(skip & 32)) NEPB_LAUNCH(lu_diag_inv_kernel, dim3(L.front_count, nb), 128, DIAG_SMEM, sd->dev, sd->fr_items.p + L.front_begin, F, lu->piv.p, lu->info.p);
    if (L.pn_count && !(skip & 4)) NEPB_LAUNCH(lu_panel_inv_kernel, dim3(L.pn_count, nb), 256, PANEL_SMEM, sd->dev, sd->pn_items.p + L.pn_begin, F, lu->piv.p);
}

static void factor_level_schur(nepb_lu* lu, int l) {
    LuSymbolicDev* sd = lu->sym;
    const int nb = lu->nb;
    double2* F = (double2*)lu->fronts.p;
    const auto& L = sd->lv[l];
    if (lu_skip_mask() & 1) return;
    // few shifts in flight: one tile per CTA (twice the CTAs, half the time per CTA); otherwise two column tiles per CTA share
    // the L21 tile
    if (L.sc_count && sd->schur_ring) {
        static const int force_one = getenv("NEPB_LU_SCHUR_ONE") ? atoi(getenv("NEPB_LU_SCHUR_ONE")) : -1;
        const bool one = force_one >= 0 ? force_one != 0 : (int64_t)nb * L.sp_count < 2 * sm_count();
        static const size_t pad = getenv("NEPB_LU_SCHUR_PAD") ? (size_t)atoi(getenv("NEPB_LU_SCHUR_PAD")) * 1024 : 0;  // occupancy experiments
        if (sd->schur_tn == 32)
            NEPB_LAUNCH(lu_schur_ring32_kernel, dim3(one ? L.sc_count : L.sp_count, nb), 128, sr_smem(32, 2) + pad, sd->dev,
                        one ? sd->sc_items.p + L.sc_begin : sd->sp_items.p + L.sp_begin, F);
        else
            NEPB_LAUNCH(lu_schur_ring_kernel, dim3(one ? L.sc_count : L.sp_count, nb), 256, sr_smem(64, 3) + pad, sd->dev,
                        one ? sd->sc_items.p + L.sc_begin : sd->sp_items.p + L.sp_begin, F);
        return;
    }
    static const bool cinit = (getenv("NEPB_LU_SCHUR_CINIT") && atoi(getenv("NEPB_LU_SCHUR_CINIT")) != 0);
    const bool one_tile = (int64_t)nb * L.sp_count < 2 * sm_count();
    if (L.sc_count && sd->schur_dmma) {
        const dim3 grid(one_tile ? L.sc_count : L.sp_count, nb);
        const int4* items = one_tile ? sd->sc_items.p + L.sc_begin : sd->sp_items.p + L.sp_begin;
        if (cinit) NEPB_LAUNCH(lu_schur_dmma_kernel<true>, grid, 256, sd->smem_schur_dmma, sd->dev, items, F);
        else NEPB_LAUNCH(lu_schur_dmma_kernel<false>, grid, 256, sd->smem_schur_dmma, sd->dev, items, F);
    }
    else if (L.sc_count && sd->schur_pipe && (int64_t)nb * L.sp_count < 2 * sm_count())
        NEPB_LAUNCH(lu_schur_pipe_kernel, dim3(L.sc_count, nb), 256, sd->smem_schur_pipe, sd->dev, sd->sc_items.p + L.sc_begin, F);
    else if (L.sc_count && sd->schur_pipe)
        NEPB_LAUNCH(lu_schur_pipe_kernel, dim3(L.sp_count, nb), 256, sd->smem_schur_pipe, sd->dev, sd->sp_items.p + L.sp_begin, F);
    else if (L.sc_count)
        NEPB_LAUNCH(lu_schur_kernel, dim3(L.sc_count, nb), 256, sd->smem_schur, sd->dev, sd->sc_items.p + L.sc_begin, F);
}

int lu_factor_device(nepb_lu* lu) {
    int rc = factor_prologue(lu);
    if (rc) return rc;
    for (int l = 0; l < lu->sym->S.nlevels; ++l) {
        factor_level_panels(lu, l);
        factor_level_schur(lu, l);
    }
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

// Solve for the shifts [shift0, shift0+nb): Bdev holds right-hand sides [b][n][k] row-major (rhs_stride = n*k) or one
// shared block (rhs_stride = 0); Xdev [b][n][k].  Asynchronous on the library stream.
// grow the solve scratch for nb shifts x k right-hand sides (must happen outside graph capture)
int lu_solve_reserve(nepb_lu* lu, int nb, int k) {
    LuSymbolicDev* sd = lu->sym;
    const LuSymbolic& S = sd->S;
    NEPB_CUDA(lu->xp.reserve((size_t)2 * nb * S.n * k));
    NEPB_CUDA(lu->w.reserve((size_t)2 * nb * S.w_total * k));
    NEPB_CUDA(lu->part.reserve((size_t)2 * nb * std::max(sd->part_slots, 1) * S.max_np * k));
    return NEPB_OK;
}

// The solve walks the tree level by level: forward bottom-up, backward top-down.  Shared memory is carved with the level's
// own largest pivot block: the many small fronts at the bottom of the tree then fit more CTAs per SM.
struct SolveCtx {
    LuSymbolicDev* sd;
    int nb, k;
    const double2* F;
    const int* piv;
    double2 *Xp, *W, *part;
    size_t smem;
    bool chain_fwd = true;  // forward chain tails by the cluster kernel (false: per level, beside the factorisation)
};

static inline size_t chain_smem_bytes(int k) { return ((size_t)32 * 33 + (size_t)2 * 32 * k) * 16; }
static const int2*& fc_ptr(LuSymbolicDev* sd, const LuLevel& L) {
    static thread_local const int2* p;
    p = sd->fc_items.p + L.fc_begin;
    return p;
}
static const int2*& bc_ptr(LuSymbolicDev* sd, const LuLevel& L) {
    static thread_local const int2* p;
    p = sd->bc_items.p + L.bc_begin;
    return p;
}
// one cluster of CHAIN_C CTAs per (chain, shift)
static void launch_chain_cluster(const void* kernel, int nchains, int nb, int k, void** args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nchains * CHAIN_C), (unsigned)nb, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = chain_smem_bytes(k);
    cfg.stream = stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CHAIN_C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelExC(&cfg, kernel, args);
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

template <int CK>
static void solve_forward_level(const SolveCtx& c, int l) {
    LuSymbolicDev* sd = c.sd;
    const auto& L = sd->lv[l];
    const size_t sml = solve_smem_bytes(L.max_np, c.k);
    // per-level kernels: without the chain tails when the chain kernel takes them (single solves), with them when the forward
    // substitution runs beside the factorisation (level l as soon as its panels are final)
    const int nfr = c.chain_fwd ? L.sfr_count : L.sfr_count_all, nfu = c.chain_fwd ? L.fu_count : L.fu_count_all;
    if (nfr)
        NEPB_LAUNCH((lu_forward_kernel<CK>), dim3(nfr, c.nb), 256, sml, sd->dev, sd->sfr_items.p + L.sfr_begin, c.F, c.piv, c.Xp, c.W, c.k,
                    L.max_np);
    if (nfu)
        NEPB_LAUNCH((lu_forward_update_kernel<CK>), dim3(nfu, c.nb), 128, sml, sd->dev, sd->fu_items.p + L.fu_begin, c.F,
                    (const double2*)c.Xp, c.W, c.k, L.max_np);
    if (L.fc_count && c.chain_fwd) {
        void* args[] = {(void*)&sd->dev, (void*)&fc_ptr(sd, L), (void*)&sd->chain_fronts.p, (void*)&c.F, (void*)&c.piv, (void*)&c.Xp, (void*)&c.W, (void*)&c.k};
        launch_chain_cluster((const void*)lu_forward_chain_cluster_kernel, L.fc_count, c.nb, c.k, args);
    }
}

// partial products of level l over the solution rows of levels >= l + 2 (everything above the parents)
template <int CK>
static void solve_backward_partials(const SolveCtx& c, int l) {
    LuSymbolicDev* sd = c.sd;
    const auto& L = sd->lv[l];
    if (L.bp_count)
        NEPB_LAUNCH((lu_backward_partial_kernel<CK>), dim3(L.bp_count, c.nb), 128, solve_smem_bytes(L.max_np, c.k), sd->dev,
                    sd->bp_items.p + L.bp_begin, c.F, (const double2*)c.Xp, c.part, sd->part_slots, L.max_np, c.k);
}

template <int CK>
static void solve_backward_level(const SolveCtx& c, int l) {
    LuSymbolicDev* sd = c.sd;
    const auto& L = sd->lv[l];
    const size_t sml = solve_smem_bytes(L.max_np, c.k);
    if (L.bc_count) {
        void* args[] = {(void*)&sd->dev, (void*)&bc_ptr(sd, L), (void*)&sd->chain_fronts.p, (void*)&c.F, (void*)&c.Xp, (void*)&c.k};
        launch_chain_cluster((const void*)lu_backward_chain_cluster_kernel, L.bc_count, c.nb, c.k, args);
    }
    if (L.sfr_count)
        NEPB_LAUNCH((lu_backward_kernel<CK>), dim3(L.sfr_count, c.nb), 256, sml, sd->dev, sd->sfr_items.p + L.sfr_begin, c.F, c.Xp,
                    (const double2*)c.part, sd->part_slots, L.max_np, c.k);
}

// right-hand-side columns held in registers per pass
static int solve_ck(int k) { return k == 1 ? 1 : k <= 4 ? 4 : k <= 8 ? 8 : (k <= 10 || (k > 16 && k <= 20)) ? 10 : 16; }
static void solve_forward_level(const SolveCtx& c, int l) {
    switch (solve_ck(c.k)) {
        case 1: solve_forward_level<1>(c, l); break;
        case 4: solve_forward_level<4>(c, l); break;
        case 8: solve_forward_level<8>(c, l); break;
        case 10: solve_forward_level<10>(c, l); break;
        default: solve_forward_level<16>(c, l); break;
    }
}
static void solve_backward_partials(const SolveCtx& c, int l) {
    switch (solve_ck(c.k)) {
        case 1: solve_backward_partials<1>(c, l); break;
        case 4: solve_backward_partials<4>(c, l); break;
        case 8: solve_backward_partials<8>(c, l); break;
        case 10: solve_backward_partials<10>(c, l); break;
        default: solve_backward_partials<16>(c, l); break;
    }
}
static bool level_has_partials(const SolveCtx& c, int l) { return c.sd->lv[l].bp_count > 0; }
static void solve_backward_level(const SolveCtx& c, int l) {
    switch (solve_ck(c.k)) {
        case 1: solve_backward_level<1>(c, l); break;
        case 4: solve_backward_level<4>(c, l); break;
        case 8: solve_backward_level<8>(c, l); break;
        case 10: solve_backward_level<10>(c, l); break;
        default: solve_backward_level<16>(c, l); break;
    }
}

static int solve_ctx(nepb_lu* lu, int shift0, int nb, int k, SolveCtx* c) {
    LuSymbolicDev* sd = lu->sym;
    const LuSymbolic& S = sd->S;
    NEPB_CHECK_ARG(shift0 >= 0 && nb >= 1 && shift0 + nb <= lu->nb, "shift window out of range");
    NEPB_CHECK_ARG(k >= 1 && k <= 256, "number of right-hand sides per solve must be in 1..256 (k=%d)", k);
    const size_t smem = solve_smem_bytes(S.max_np, k);
    NEPB_CHECK_ARG(smem <= 200 * 1024, "k=%d right-hand sides with %d pivot columns per front exceed shared memory", k, S.max_np);
    int rc = lu_solve_reserve(lu, nb, k);
    if (rc) return rc;
    c->sd = sd;
    c->nb = nb;
    c->k = k;
    c->F = (const double2*)lu->fronts.p + (size_t)shift0 * S.front_total;
    c->piv = lu->piv.p + (size_t)shift0 * S.n;
    c->Xp = (double2*)lu->xp.p;
    c->W = (double2*)lu->w.p;
    c->part = (double2*)lu->part.p;
    c->smem = smem;
    return NEPB_OK;
}

int lu_solve_device(nepb_lu* lu, int shift0, int nb, int k, const double2* Bdev, size_t rhs_stride, double2* Xdev) {
    SolveCtx c;
    int rc = solve_ctx(lu, shift0, nb, k, &c);
    if (rc) return rc;
    const int n = c.sd->S.n, nlev = c.sd->S.nlevels;
    dim3 pg((unsigned)(((size_t)n * k + 255) / 256), nb);
    NEPB_LAUNCH(lu_permute_in_kernel, pg, 256, 0, n, k, c.sd->rperm.p, c.sd->dr.p, Bdev, rhs_stride, c.Xp);
    for (int l = 0; l < nlev; ++l) solve_forward_level(c, l);
    for (int l = nlev - 1; l >= 0; --l) {
        solve_backward_partials(c, l);
        solve_backward_level(c, l);
    }
    NEPB_LAUNCH(lu_permute_out_kernel, pg, 256, 0, n, k, c.sd->iperm.p, c.sd->dc.p, c.Xp, Xdev, (size_t)n * k);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

// Factorise all lu->nb shifts and solve them against Bdev in one pipelined launch sequence: the forward substitution
// walks the tree in the same bottom-up order as the factorisation, so level l of the forward solve is enqueued on a
// second stream as soon as the panels of level l are final and runs beside the Schur update / the next levels of the
// factorisation.  Only the backward substitution remains on the critical path after the root is factorised.
// `ev` holds at least 3 * nlevels + 2 events (no timing).  Works both eagerly and under stream capture (fork / join).
int lu_factor_solve_pipelined(nepb_lu* lu, int k, const double2* Bdev, size_t rhs_stride, double2* Xdev, cudaStream_t side,
                              cudaEvent_t* ev) {
    SolveCtx c;
    int rc = solve_ctx(lu, 0, lu->nb, k, &c);
    if (rc) return rc;
    const int n = c.sd->S.n, nlev = c.sd->S.nlevels, nb = lu->nb;
    c.chain_fwd = false;
    cudaStream_t s0 = stream();
    dim3 pg((unsigned)(((size_t)n * k + 255) / 256), nb);
    NEPB_CUDA(cudaEventRecord(ev[0], s0));  // fork: the side stream starts after everything already queued on s0
    NEPB_CUDA(cudaStreamWaitEvent(side, ev[0], 0));
    set_current_stream(side);
    NEPB_LAUNCH(lu_permute_in_kernel, pg, 256, 0, n, k, c.sd->rperm.p, c.sd->dr.p, Bdev, rhs_stride, c.Xp);
    set_current_stream(s0);
    rc = factor_prologue(lu);
    if (rc) return rc;
    for (int l = 0; l < nlev; ++l) {
        factor_level_panels(lu, l);
        NEPB_CUDA(cudaEventRecord(ev[l + 1], s0));
        factor_level_schur(lu, l);
        NEPB_CUDA(cudaStreamWaitEvent(side, ev[l + 1], 0));
        set_current_stream(side);
        if (!(lu_skip_mask() & 8)) solve_forward_level(c, l);
        set_current_stream(s0);
    }
    NEPB_CUDA(cudaEventRecord(ev[nlev + 1], side));  // join
    NEPB_CUDA(cudaStreamWaitEvent(s0, ev[nlev + 1], 0));
    // backward, top-down: the partial products of level l need the solution rows of the levels >= l + 2 only, so they run on
    // the side stream beside the fronts of level l + 1; per level the critical path is one launch
    cudaEvent_t* evB = ev + nlev + 2;      // evB[l]: fronts of level l solved
    cudaEvent_t* evP = ev + 2 * nlev + 2;  // evP[l]: partial products of level l formed
    for (int l = nlev - 1; l >= 0; --l) {
        const bool hp = level_has_partials(c, l);
        if (hp) {
            if (l + 2 < nlev) NEPB_CUDA(cudaStreamWaitEvent(side, evB[l + 2], 0));
            set_current_stream(side);
            if (!(lu_skip_mask() & 16)) solve_backward_partials(c, l);
            set_current_stream(s0);
            NEPB_CUDA(cudaEventRecord(evP[l], side));
            NEPB_CUDA(cudaStreamWaitEvent(s0, evP[l], 0));
        }
        if (!(lu_skip_mask() & 16)) solve_backward_level(c, l);
        NEPB_CUDA(cudaEventRecord(evB[l], s0));
    }
    NEPB_LAUNCH(lu_permute_out_kernel, pg, 256, 0, n, k, c.sd->iperm.p, c.sd->dc.p, c.Xp, Xdev, (size_t)n * k);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

// lu->sol = M(sigma_shift)^-1 lu->rhs for k right-hand sides.  The ~4 launches per tree level are captured once per
// (shift, k) into a CUDA graph and replayed: the single-shift solve of iar / tiar / resinv is launch-bound otherwise.
int lu_solve_staged(nepb_lu* lu, int shift, int k) {
    static const bool use_graph = !(getenv("NEPB_SOLVE_GRAPH") && atoi(getenv("NEPB_SOLVE_GRAPH")) == 0);
    int rc = lu_solve_reserve(lu, 1, k);
    if (rc) return rc;
    if (!use_graph) return lu_solve_device(lu, shift, 1, k, (const double2*)lu->rhs.p, 0, (double2*)lu->sol.p);
    const auto key = std::make_pair(shift, k);
    auto it = lu->solve_graphs.find(key);
    if (it == lu->solve_graphs.end() || it->second.rhs != lu->rhs.p || it->second.sol != lu->sol.p) {
        if (it != lu->solve_graphs.end()) {
            cudaGraphExecDestroy(it->second.exec);
            lu->solve_graphs.erase(it);
        }
        const int64_t l0 = g_launches.load();
        cudaGraph_t graph = nullptr;
        NEPB_CUDA(cudaStreamBeginCapture(stream(), cudaStreamCaptureModeThreadLocal));
        rc = lu_solve_device(lu, shift, 1, k, (const double2*)lu->rhs.p, 0, (double2*)lu->sol.p);
        cudaError_t e = cudaStreamEndCapture(stream(), &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        NEPB_CUDA(e);
        nepb_lu::SolveGraph g;
        g.kernels = (int)(g_launches.load() - l0);
        g_launches.fetch_sub(g.kernels);
        g.rhs = lu->rhs.p;
        g.sol = lu->sol.p;
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        NEPB_CUDA(e);
        it = lu->solve_graphs.emplace(key, g).first;
    }
    NEPB_CUDA(cudaGraphLaunch(it->second.exec, stream()));
    g_launches.fetch_add(it->second.kernels);
    return NEPB_OK;
}

int spmf_apply_device(const nepb_spmf* h, int mode, int k, int q, const double2* dV, const double* C, double2* dZ);
int upload_colmajor(int64_t n, int kc, const double* host, int64_t ld, DevBuf<double>& stage, double* dst, int ldd, int k0);
int download_colmajor(int64_t n, int kc, const double* src, int lds, int k0, DevBuf<double>& stage, double* host, int64_t ld);

int lu_fetch_info(nepb_lu* lu) {
    lu->h_info.resize(lu->nb);
    NEPB_CUDA(cudaMemcpyAsync(lu->h_info.data(), lu->info.p, sizeof(LuInfo) * lu->nb, cudaMemcpyDeviceToHost, stream()));
    NEPB_CUDA(cudaStreamSynchronize(stream()));
    return NEPB_OK;
}

int lu_refactor(nepb_lu* lu, int nshift, const double* coef) {
    NEPB_CHECK_ARG(nshift >= 1 && nshift <= lu->cap, "refactor: %d shifts exceed the capacity %d of this handle", nshift, lu->cap);
    lu->nb = nshift;
    lu->h_coef.assign(coef, coef + (size_t)2 * nshift * lu->op->p);
    NEPB_CUDA(cudaStreamSynchronize(stream()));  // h_coef / h_info may still be read by an earlier async copy
    NEPB_CUDA(cudaMemcpyAsync(lu->coef.p, lu->h_coef.data(), sizeof(double) * lu->h_coef.size(), cudaMemcpyHostToDevice, stream()));
    return lu_factor_device(lu);
}

int lu_create(const nepb_spmf* h, int nshift, const double* coef, nepb_lu** out) {
    *out = nullptr;
    LuSymbolicDev* sd = nullptr;
    int rc = lu_symbolic_get(h, &sd);
    if (rc) return rc;
    nepb_lu* lu = new nepb_lu();
    lu->op = h;
    lu->sym = sd;
    lu->nb = lu->cap = nshift;
    const LuSymbolic& S = sd->S;
    cudaError_t e = lu->fronts.alloc((size_t)2 * nshift * S.front_total);
    if (e == cudaSuccess) e = lu->piv.alloc((size_t)nshift * S.n);
    if (e == cudaSuccess) e = lu->info.alloc(nshift);
    if (e == cudaSuccess) e = lu->coef.alloc((size_t)2 * nshift * h->p);
    if (e != cudaSuccess) {
        set_error("allocating %d factorisations (%.1f MB each) failed: %s", nshift, S.front_total * 16e-6, cudaGetErrorString(e));
        delete lu;
        return e == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;
    }
    rc = lu_refactor(lu, nshift, coef);
    if (!rc) rc = lu_fetch_info(lu);
    // Static-pivoting fallback (NEPB_LU_MATCHING: 0 = never, 1 = when the plain factorisation met a zero or tiny pivot
    // [default], 2 = always): permute / scale with the maximum-product matching of the first offending shift and factorise again
    static const int matching = getenv("NEPB_LU_MATCHING") ? atoi(getenv("NEPB_LU_MATCHING")) : 1;
    if (!rc && matching > 0) {
        int bad = -1;
        for (int b = 0; b < nshift && bad < 0; ++b)
            if (lu_info_suspicious(lu->h_info[b])) bad = b;
        if (matching >= 2 && bad < 0 && !sd->matched()) bad = 0;
        if (bad >= 0 && sd->matched() && !(lu->h_info[bad].flags & 3) && lu->h_info[bad].nperturbed == 0)
            bad = -1;  // small pivots under an existing matching: genuinely close to singular, the solves verify themselves
        if (bad >= 0) {
            LuSymbolicDev* md = nullptr;
            rc = lu_symbolic_make_matched(h, coef + (size_t)2 * bad * h->p, &md);
            if (!rc) {
                lu->sym = md;
                for (auto& kv : lu->solve_graphs) cudaGraphExecDestroy(kv.second.exec);
                lu->solve_graphs.clear();
                e = lu->fronts.alloc((size_t)2 * nshift * md->S.front_total);
                if (e != cudaSuccess) {
                    set_error("allocating %d factorisations (%.1f MB each) failed: %s", nshift, md->S.front_total * 16e-6, cudaGetErrorString(e));
                    rc = e == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;
                }
            }
            if (!rc) rc = lu_refactor(lu, nshift, coef);
            if (!rc) rc = lu_fetch_info(lu);
        }
    }
    if (rc) {
        delete lu;
        return rc;
    }
    *out = lu;
    return NEPB_OK;
}

}  // namespace nepb

using namespace nepb;

extern "C" {

int nepb_lu_set_options(nepb_spmf* h, int ordering, int relax_leaf, int max_np, const int64_t* user_perm) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    NEPB_CHECK_ARG(!h->lu_symbolic, "the symbolic analysis of this operator already exists; set options before the first factorisation");
    h->lu_opt = LuOptions();
    if (ordering >= 0) h->lu_opt.ordering = ordering;
    if (relax_leaf > 0) h->lu_opt.relax_leaf = relax_leaf;
    if (max_np > 0) h->lu_opt.max_np = max_np;
    h->lu_opt_set = true;
    h->lu_user_perm.clear();
    if (user_perm) {
        h->lu_user_perm.resize(h->n);
        for (int64_t i = 0; i < h->n; ++i) h->lu_user_perm[i] = (int32_t)(user_perm[i] - h->index_base);
    }
    return NEPB_OK;
}

int nepb_lu_symbolic_info(const nepb_spmf* h, int64_t* nnz_factor, int64_t* front_entries, int* nfronts, int* nlevels, int* max_front,
                          double* flops) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    LuSymbolicDev* sd = nullptr;
    int rc = lu_symbolic_get(h, &sd);
    if (rc) return rc;
    if (nnz_factor) *nnz_factor = sd->S.nnz_factor;
    if (front_entries) *front_entries = sd->S.front_total;
    if (nfronts) *nfronts = sd->S.nsuper;
    if (nlevels) *nlevels = sd->S.nlevels;
    if (max_front) *max_front = sd->S.max_nf;
    if (flops) *flops = sd->S.flops;
    return NEPB_OK;
}

int nepb_lu_symbolic_get(const nepb_spmf* h, int32_t* perm, int32_t* parent, int32_t* sn_ptr, int32_t* sn_parent, int32_t* sn_rows,
                         int32_t* sn_level) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    LuSymbolicDev* sd = nullptr;
    int rc = lu_symbolic_get(h, &sd);
    if (rc) return rc;
    const LuSymbolic& S = sd->S;
    if (perm) memcpy(perm, S.perm.data(), sizeof(int32_t) * S.n);
    if (parent) memcpy(parent, S.parent.data(), sizeof(int32_t) * S.n);
    if (sn_ptr) memcpy(sn_ptr, S.sn_ptr.data(), sizeof(int32_t) * (S.nsuper + 1));
    if (sn_parent) memcpy(sn_parent, S.sn_parent.data(), sizeof(int32_t) * S.nsuper);
    if (sn_rows)
        for (int s = 0; s < S.nsuper; ++s) sn_rows[s] = (int32_t)(S.row_ptr[s + 1] - S.row_ptr[s]);
    if (sn_level) memcpy(sn_level, S.level.data(), sizeof(int32_t) * S.nsuper);
    return NEPB_OK;
}

int nepb_lu_analyse_pattern(int64_t n, const int64_t* colptr, const int64_t* rowval, int index_base, int ordering, int relax_leaf,
                            int max_np, int32_t* perm, int32_t* parent, int32_t* colcount, double* stats, int32_t* sn_ptr,
                            int32_t* sn_rows, int32_t* sn_level) {
    NEPB_CHECK_ARG(n >= 1 && n < ((int64_t)1 << 31) && colptr && rowval, "bad arguments");
    const int64_t nnz = colptr[n] - index_base;
    NEPB_CHECK_ARG(nnz >= 0 && nnz < ((int64_t)1 << 31), "pattern too large");
    std::vector<int32_t> rp(n + 1), ci(nnz);
    for (int64_t j = 0; j <= n; ++j) rp[j] = (int32_t)(colptr[j] - index_base);
    for (int64_t e = 0; e < nnz; ++e) {
        ci[e] = (int32_t)(rowval[e] - index_base);
        NEPB_CHECK_ARG(ci[e] >= 0 && ci[e] < n, "index out of range");
    }
    LuOptions opt;
    if (ordering >= 0) opt.ordering = ordering;
    if (relax_leaf > 0) opt.relax_leaf = relax_leaf;
    if (max_np > 0) opt.max_np = std::min(32, max_np);
    LuSymbolic S;
    // the analysis symmetrises the pattern, so CSC and CSR input are equivalent
    int rc = lu_symbolic_analyse((int)n, rp.data(), ci.data(), nullptr, opt, S);
    if (rc) return rc;
    if (perm) memcpy(perm, S.perm.data(), sizeof(int32_t) * n);
    if (parent) memcpy(parent, S.parent.data(), sizeof(int32_t) * n);
    if (colcount) memcpy(colcount, S.colcount.data(), sizeof(int32_t) * n);
    if (stats) {
        stats[0] = (double)S.nnz_factor;
        stats[1] = (double)S.front_total;
        stats[2] = S.nsuper;
        stats[3] = S.nlevels;
        stats[4] = S.max_nf;
        stats[5] = S.max_np;
        stats[6] = S.flops;
        stats[7] = (double)S.w_total;
    }
    if (sn_ptr) memcpy(sn_ptr, S.sn_ptr.data(), sizeof(int32_t) * (S.nsuper + 1));
    if (sn_rows)
        for (int s = 0; s < S.nsuper; ++s) sn_rows[s] = (int32_t)(S.row_ptr[s + 1] - S.row_ptr[s]);
    if (sn_level) memcpy(sn_level, S.level.data(), sizeof(int32_t) * S.nsuper);
    return NEPB_OK;
}

int nepb_lu_matching(int64_t n, const int64_t* colptr, const int64_t* rowval, int index_base, const double* absval,
                     int32_t* row_of_col, double* dr, double* dc) {
    NEPB_CHECK_ARG(n >= 1 && n < ((int64_t)1 << 31) && colptr && rowval && absval && row_of_col, "bad arguments");
    const int64_t nnz = colptr[n] - index_base;
    NEPB_CHECK_ARG(nnz >= 0 && nnz < ((int64_t)1 << 31), "pattern too large");
    std::vector<int32_t> cp(n + 1), ri((size_t)nnz), match;
    for (int64_t j = 0; j <= n; ++j) cp[j] = (int32_t)(colptr[j] - index_base);
    for (int64_t e = 0; e < nnz; ++e) {
        ri[e] = (int32_t)(rowval[e] - index_base);
        NEPB_CHECK_ARG(ri[e] >= 0 && ri[e] < n, "row index out of range");
    }
    // the matching routine is orientation-agnostic: handing it the CSC arrays matches columns (outer) to rows (inner)
    std::vector<double> douter, dinner;
    const int matched = max_product_matching((int)n, cp.data(), ri.data(), absval, match, douter, dinner);
    if (matched < n) {
        set_error("structurally singular: only %d of %lld columns can be matched to a nonzero", matched, (long long)n);
        return NEPB_E_SINGULAR;
    }
    memcpy(row_of_col, match.data(), sizeof(int32_t) * n);
    if (dc) memcpy(dc, douter.data(), sizeof(double) * n);
    if (dr) memcpy(dr, dinner.data(), sizeof(double) * n);
    return NEPB_OK;
}

int nepb_lu_create(const nepb_spmf* h, int nshift, const double* coef, nepb_lu** out) {
    NEPB_CHECK_ARG(h && coef && out, "NULL argument");
    NEPB_CHECK_ARG(nshift >= 1 && nshift <= 65535, "nshift must be in 1..65535");
    return lu_create(h, nshift, coef, out);
}

int nepb_lu_destroy(nepb_lu* lu) {
    delete lu;
    return NEPB_OK;
}

int nepb_lu_status(const nepb_lu* lu, int shift, int* flags, int* nperturbed, double* min_pivot_ratio) {
    NEPB_CHECK_ARG(lu && shift >= 0 && shift < lu->nb, "bad arguments");
    const LuInfo& I = lu->h_info[shift];
    if (flags) *flags = I.flags | (lu->sym->matched() ? 8 : 0);
    if (nperturbed) *nperturbed = I.nperturbed;
    if (min_pivot_ratio) {
        double r;
        memcpy(&r, &I.minpiv_bits, 8);
        *min_pivot_ratio = r;
    }
    return NEPB_OK;
}

}  // extern "C"

namespace nepb {
// Iterative refinement of lu->sol against lu->rhs (k columns, both staged on the device) with the fused SpMM residual --
// UMFPACK's control[8] in the reference (LinSolvers.jl:117-120).  Stops once the normwise backward error
// max_c |r_c|_inf / (|M|_max |x_c|_inf + |b_c|_inf) is at rounding level or no longer halves.  A factorisation that had to
// replace zero / tiny pivots is always checked, and a solve whose backward error stays above 1e-9 is an error
// (LinearAlgebra.SingularException in the reference), never a silent result.
int lu_refine_staged(nepb_lu* lu, int shift, int k, int refine_steps, bool want_berr, double* berr_out) {
    const nepb_spmf* h = lu->op;
    const int64_t n = h->n;
    const LuInfo& I = lu->h_info[shift];
    const bool suspicious = lu_info_suspicious(I);
    *berr_out = 0.0;
    if (refine_steps <= 0 && !want_berr && !suspicious) return NEPB_OK;
    NEPB_CUDA(lu->res.reserve((size_t)2 * n * k));
    NEPB_CUDA(lu->cor.reserve((size_t)2 * n * k));
    NEPB_CUDA(lu->colmax.reserve(3 * 256));
    const double* coef = lu->h_coef.data() + (size_t)2 * shift * h->p;
    const size_t cnt = (size_t)n * k;
    const unsigned gb = (unsigned)((cnt + 255) / 256);
    double amax;
    memcpy(&amax, &I.amax_plain_bits, 8);  // |M|_max of the operator itself (amax_bits is the scaled one under static pivoting)
    double prev = INFINITY, berr = 0.0;
    int rc;
    for (int it = 0;; ++it) {
        rc = spmf_apply_device(h, NEPB_COEF_SCALAR, k, k, (const double2*)lu->sol.p, coef, (double2*)lu->res.p);
        if (rc) return rc;
        NEPB_LAUNCH(residual_kernel, gb, 256, 0, cnt, (const double2*)lu->rhs.p, (double2*)lu->res.p);
        NEPB_CUDA(cudaMemsetAsync(lu->colmax.p, 0, sizeof(unsigned long long) * 3 * 256, stream()));
        NEPB_LAUNCH(colmax_kernel, gb, 256, 0, (int)n, k, (const double2*)lu->res.p, lu->colmax.p);
        NEPB_LAUNCH(colmax_kernel, gb, 256, 0, (int)n, k, (const double2*)lu->sol.p, lu->colmax.p + 256);
        NEPB_LAUNCH(colmax_kernel, gb, 256, 0, (int)n, k, (const double2*)lu->rhs.p, lu->colmax.p + 512);
        NEPB_LAUNCH_CHECK();
        unsigned long long hm[3 * 256];
        NEPB_CUDA(cudaMemcpyAsync(hm, lu->colmax.p, sizeof(hm), cudaMemcpyDeviceToHost, stream()));
        NEPB_CUDA(cudaStreamSynchronize(stream()));
        berr = 0.0;
        for (int c = 0; c < k; ++c) {
            double r, x, bb;
            memcpy(&r, &hm[c], 8);
            memcpy(&x, &hm[256 + c], 8);
            memcpy(&bb, &hm[512 + c], 8);
            const double den = amax * x + bb;
            berr = std::max(berr, den > 0 ? r / den : (r > 0 ? INFINITY : 0.0));
        }
        if (it >= refine_steps || berr <= 1.2e-16 || berr >= 0.5 * prev) break;
        prev = berr;
        rc = lu_solve_device(lu, shift, 1, k, (const double2*)lu->res.p, 0, (double2*)lu->cor.p);
        if (rc) return rc;
        NEPB_LAUNCH(axpy_kernel, gb, 256, 0, cnt, (const double2*)lu->cor.p, (double2*)lu->sol.p);
    }
    *berr_out = berr;
    if (!(berr == berr) || berr == INFINITY) {
        set_error("solution of shift %d is not finite (singular matrix?)", shift);
        return NEPB_E_SINGULAR;
    }
    if ((refine_steps > 0 || suspicious) && berr > 1e-9) {
        set_error("solve with M(sigma_%d) is inaccurate: backward error %.2e after refinement (%d pivots replaced%s); the matrix is "
                  "singular to working precision or the static pivoting failed", shift, berr, I.nperturbed,
                  lu->sym->matched() ? ", row matching in use" : "");
        return NEPB_E_SINGULAR;
    }
    return NEPB_OK;
}
}  // namespace nepb

extern "C" {

// Solve M(sigma_shift) X = B on the host interface (lin_solve, LinSolvers.jl:135-137,157-159): B, X are n x nrhs column-major.
// refine_steps > 0: iterative refinement with the fused SpMM residual (UMFPACK's control[8] in the reference); it stops
// early once the normwise backward error max_c |r_c|_inf / (|M|_max |x_c|_inf + |b_c|_inf) is at rounding level or no
// longer halves.  berr_out (optional) receives the final backward error (computed only when asked for or refining).
int nepb_lu_solve(nepb_lu* lu, int shift, int nrhs, const double* B, int64_t ldb, double* X, int64_t ldx, int refine_steps,
                  double* berr_out) {
    NEPB_CHECK_ARG(lu && B && X, "NULL argument");
    NEPB_CHECK_ARG(shift >= 0 && shift < lu->nb, "shift index %d out of range (0..%d)", shift, lu->nb - 1);
    const nepb_spmf* h = lu->op;
    const int64_t n = h->n;
    NEPB_CHECK_ARG(nrhs >= 1 && ldb >= n && ldx >= n, "bad right-hand side shape");
    if (lu->h_info[shift].flags & 3) {
        set_error("%s pivot in the factorisation of shift %d: the matrix is singular to working precision",
                  (lu->h_info[shift].flags & 2) ? "non-finite" : "zero", shift);
        return NEPB_E_SINGULAR;
    }
    const int KC = 64;  // right-hand sides per pass
    double worst = 0.0;
    for (int c0 = 0; c0 < nrhs; c0 += KC) {
        const int k = std::min(KC, nrhs - c0);
        NEPB_CUDA(lu->rhs.reserve((size_t)2 * n * k));
        NEPB_CUDA(lu->sol.reserve((size_t)2 * n * k));
        int rc = upload_colmajor(n, k, B + 2 * (size_t)c0 * ldb, ldb, lu->stage, lu->rhs.p, k, 0);
        if (rc) return rc;
        rc = lu_solve_staged(lu, shift, k);
        if (rc) return rc;
        double berr = 0.0;
        rc = lu_refine_staged(lu, shift, k, refine_steps, berr_out != nullptr, &berr);
        if (rc) return rc;
        worst = std::max(worst, berr);
        rc = download_colmajor(n, k, lu->sol.p, k, 0, lu->stage, X + 2 * (size_t)c0 * ldx, ldx);
        if (rc) return rc;
    }
    if (berr_out) *berr_out = worst;
    return NEPB_OK;
}

// Device-resident lin_solve: X[:, xcol0 : xcol0+nrhs) = alpha * M(sigma_shift)^-1 B[:, bcol0 : bcol0+nrhs); nothing crosses PCIe.
int nepb_lu_solve_block_ex(nepb_lu* lu, int shift, const nepb_block* B, int bcol0, int nrhs, nepb_block* X, int xcol0, const double* alpha,
                           int refine_steps, double* berr_out) {
    NEPB_CHECK_ARG(lu && B && X, "NULL argument");
    NEPB_CHECK_ARG(shift >= 0 && shift < lu->nb, "shift index out of range");
    const int64_t n = lu->op->n;
    NEPB_CHECK_ARG(B->n == n && X->n == n, "block row count differs from the operator size");
    NEPB_CHECK_ARG(nrhs >= 1 && nrhs <= 64 && bcol0 >= 0 && bcol0 + nrhs <= B->k && xcol0 >= 0 && xcol0 + nrhs <= X->k, "bad column windows");
    if (lu->h_info[shift].flags & 3) {
        set_error("%s pivot in the factorisation of shift %d: the matrix is singular to working precision",
                  (lu->h_info[shift].flags & 2) ? "non-finite" : "zero", shift);
        return NEPB_E_SINGULAR;
    }
    NEPB_CUDA(lu->rhs.reserve((size_t)2 * n * nrhs));
    NEPB_CUDA(lu->sol.reserve((size_t)2 * n * nrhs));
    const unsigned gb = (unsigned)((n * nrhs + 255) / 256);
    NEPB_LAUNCH(cols_gather_kernel, gb, 256, 0, n, nrhs, (const double2*)B->d.p + bcol0, B->k, (double2*)lu->rhs.p);
    int rc = lu_solve_staged(lu, shift, nrhs);
    if (rc) return rc;
    double berr = 0.0;
    rc = lu_refine_staged(lu, shift, nrhs, refine_steps, berr_out != nullptr, &berr);
    if (rc) return rc;
    if (berr_out) *berr_out = berr;
    const double2 a = alpha ? make_double2(alpha[0], alpha[1]) : make_double2(1.0, 0.0);
    NEPB_LAUNCH(cols_scatter_kernel, gb, 256, 0, n, nrhs, (const double2*)lu->sol.p, (double2*)X->d.p + xcol0, X->k, a);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

int nepb_lu_solve_block(nepb_lu* lu, int shift, const nepb_block* B, int bcol0, int nrhs, nepb_block* X, int xcol0, const double* alpha) {
    return nepb_lu_solve_block_ex(lu, shift, B, bcol0, nrhs, X, xcol0, alpha, 0, nullptr);
}

}  // extern "C"
