// Fused multi-term CSR SpMM for SPMF operators on sm_100a, and the C ABI around it.
//
//   Z(n x q) = sum_{i<p} A_i * (V(n x k) * C_i(k x q))
//
// All p matrices share one union pattern: one int32 column index per nonzero and the p values
// interleaved right behind each other (vals[nz][p]), so the kernel walks every A_i in ONE pass over
// the pattern and the coefficient combine happens in registers.  V / Z live in HBM row-major
// (k*16 contiguous bytes per row) so a gathered row V[col,:] is one coalesced read.
//
// Replaces (reference, relative to src/): NEPTypes.jl:276-319 (compute_MM), :322-367 (compute_Mder),
// :972-1011 and :1130-1160 (compute_Mlincomb), types_poly.jl:44-76, method_nleigs.jl:456-472 (BBCC*z),
// errmeasure.jl:128-130 (one SpMV per Ritz pair -> one multi-lambda SpMM).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "common.h"
#include "spmf_host.h"

namespace nepb {

// ---------------------------------------------------------------------------------------------
// load helpers: matrix stream bypasses L1 (read once), V gathers allocate in L1 (stencil reuse)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld_stream_f64x2(const double* p) {
    double2 r;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
    double r;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream_s32(const int* p) {
    int r;
    asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

template <int VW>
__device__ __forceinline__ void load_vals(const double* __restrict__ vals, size_t idx, double (&v)[VW]) {
    const double* p = vals + idx * VW;
    if constexpr (VW % 2 == 0) {
#pragma unroll
        for (int t = 0; t < VW; t += 2) {
            double2 x = ld_stream_f64x2(p + t);
            v[t] = x.x;
            v[t + 1] = x.y;
        }
    } else {
#pragma unroll
        for (int t = 0; t < VW; ++t) v[t] = ld_stream_f64(p + t);
    }
}

constexpr int MAXP = 16;
struct CoefP {
    double2 c[MAXP];
};

__device__ __forceinline__ void cfma(double2& acc, const double2 a, const double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

// m = sum_i c_i * val_i for one nonzero; CA: values complex (VW = 2p) else real (VW = p)
template <int VW, bool CA, class CF>
__device__ __forceinline__ double2 combine(const double (&v)[VW], CF&& cf) {
    double2 m = make_double2(0.0, 0.0);
    if constexpr (CA) {
#pragma unroll
        for (int i = 0; i < VW / 2; ++i) cfma(m, cf(i), make_double2(v[2 * i], v[2 * i + 1]));
    } else {
#pragma unroll
        for (int i = 0; i < VW; ++i) {
            double2 c = cf(i);
            m.x = fma(c.x, v[i], m.x);
            m.y = fma(c.y, v[i], m.y);
        }
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// K1/K2: SCALAR (C_i = c_i I) and DIAG (C_i = diag c_i[.]) modes.
// A row is owned by G = GC*GN lanes: GC lanes split the dense columns (cyclically, CPT each, so one
// gather instruction reads GC*16 contiguous bytes), GN lanes split the row's nonzeros (U-deep
// unrolled so U*GN independent loads are in flight per row) and are reduced with warp shuffles.
// ---------------------------------------------------------------------------------------------
template <int VW, bool CA, bool DIAG, int GC, int GN, int CPT, int U>
__global__ void __launch_bounds__(256) spmm_fused_kernel(int n, int kt, int ldv, int ldz, const int* __restrict__ rowptr,
                                                         const int* __restrict__ colind, const double* __restrict__ vals,
                                                         const double2* __restrict__ V, double2* __restrict__ Z,
                                                         const CoefP cp, const double2* __restrict__ cdiag, int p) {
    constexpr int G = GC * GN;
    constexpr int ROWS = 256 / G;
    const int tid = threadIdx.x;
    const int g = tid % G;
    const int gc = g % GC;
    const int gn = g / GC;
    const int row = blockIdx.x * ROWS + tid / G;

    // DIAG: per-column coefficients of this lane's CPT columns, kept in registers
    double2 cd[DIAG ? CPT : 1][DIAG ? (CA ? VW / 2 : VW) : 1];
    if constexpr (DIAG) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
#pragma unroll
            for (int i = 0; i < (CA ? VW / 2 : VW); ++i) cd[j][i] = (c < kt) ? cdiag[i + p * c] : make_double2(0.0, 0.0);
        }
    }

    double2 acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = make_double2(0.0, 0.0);

    int start = 0, end = 0;
    if (row < n) {
        start = rowptr[row];
        end = rowptr[row + 1];
    }
    for (int base = start; base < end; base += GN * U) {
        int col[U];
        double v[U][VW];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * GN + gn;
            ok[u] = idx < end;
            col[u] = 0;
            if (ok[u]) {
                col[u] = ld_stream_s32(colind + idx);
                load_vals<VW>(vals, (size_t)idx, v[u]);
            }
        }
        double2 x[U][CPT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int c = gc + j * GC;
                x[u][j] = make_double2(0.0, 0.0);
                if (ok[u] && c < kt) x[u][j] = __ldg(V + (size_t)col[u] * ldv + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ok[u]) {
                if constexpr (!DIAG) {
                    const double2 m = combine<VW, CA>(v[u], [&](int i) { return cp.c[i]; });
#pragma unroll
                    for (int j = 0; j < CPT; ++j) cfma(acc[j], m, x[u][j]);
                } else {
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const double2 m = combine<VW, CA>(v[u], [&](int i) { return cd[j][i]; });
                        cfma(acc[j], m, x[u][j]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int off = GC; off < G; off <<= 1) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, off);
            acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, off);
        }
    }
    if (row < n && gn == 0) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
            if (c < kt) Z[(size_t)row * ldz + c] = acc[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1t: the same product for 5..32 dense columns with the rows of V a tile of matrix rows needs staged in shared memory.
// With k columns a gathered row of V is k*16 bytes and every nonzero gathers one: the gathers (not the matrix stream) are
// what the memory system sees, 21 x 128 B per row of the C4 stencil at k = 8.  Neighbouring rows of a sparse matrix share
// most of their columns, so the operator is cut once (host, spmf_build_tiles) into tiles of <= R consecutive rows with
// <= 6R distinct columns.  A CTA
//   1. issues asynchronous 16-byte copies (cp.async: no registers held, deep memory-level parallelism) of the tile's
//      distinct V rows into shared memory,
//   2. streams the tile's slice of the interleaved matrix values once, coalesced, and reduces every nonzero to its
//      combined coefficient m = sum_i c_i a_i (SCALAR mode; DIAG keeps the p raw values), stored next to the 16-bit
//      tile-local column index,
//   3. after one wait runs the row products out of shared memory: 8 lanes per row, and those 8 lanes read 128 contiguous
//      bytes of a V row per instruction -- one conflict-free shared-memory wavefront per nonzero and 8 columns.
// Other resident CTAs are in their copy phase meanwhile.  The bound is the shared-memory pipe (one wavefront per clock
// and SM), not HBM: see DESIGN.md.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src));
}

// BULK = true (NEPB_SPMM_BULK=1; compiled and inspected, not yet measured -- see DESIGN.md) stages every V row with ONE
// TMA bulk copy (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier) instead of kt 16-byte LDGSTS
// copies: the copies then bypass the LSU pipe this kernel is bound by.
template <int VW, bool CA, bool DIAG, int CPT, bool BULK>
__global__ void __launch_bounds__(256) spmm_tiled_kernel(int kt, unsigned kinv, int ldv, int ldz, int max_cols, int max_nnz,
                                                         const int4* __restrict__ tiles, const int* __restrict__ tile_cols,
                                                         const int* __restrict__ rowptr, const uint16_t* __restrict__ lidx,
                                                         const double* __restrict__ vals, const double2* __restrict__ V,
                                                         double2* __restrict__ Z, const CoefP cp, const double2* __restrict__ cdiag, int p) {
    constexpr int GC = 8;                 // lanes per row; lane gc owns the dense columns gc, gc + 8, ..
    constexpr int MW = DIAG ? VW : 2;     // doubles kept per nonzero: raw term values (DIAG) or the combined coefficient
    extern __shared__ double2 sV[];                          // [distinct columns][kt]
    double* sM = (double*)(sV + (size_t)max_cols * kt);      // [tile nonzeros][MW]
    uint16_t* sL = (uint16_t*)(sM + (size_t)max_nnz * MW);   // [tile nonzeros] tile-local column
    int* sRp = (int*)(sL + ((max_nnz + 7) & ~7));            // [rows + 1] row pointers relative to the tile
    const int4 t0 = tiles[2 * blockIdx.x], t1 = tiles[2 * blockIdx.x + 1];
    const int row0 = t0.x, nrows = t0.y, ncols = t0.w, nz0 = t1.x, nnz = t1.y;
    const int tid = threadIdx.x, nth = blockDim.x;
    unsigned bar_s = 0;
    if constexpr (BULK) {  // 1. the tile's rows of V, one bulk copy per row
        bar_s = (unsigned)__cvta_generic_to_shared(sRp + 36);  // 8-byte mbarrier behind the row pointers
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::);  // the async proxy must see the initialised barrier
        }
        __syncthreads();
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(ncols * kt * 16) : "memory");
        const int* cols = tile_cols + t0.z;
        for (int dcol = tid; dcol < ncols; dcol += nth) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&sV[dcol * kt]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(V + (size_t)cols[dcol] * ldv), "r"(kt * 16), "r"(bar_s)
                         : "memory");
        }
    } else {  // 1. the tile's rows of V: consecutive threads copy consecutive 16-byte pieces of one row
        const int total = ncols * kt;
        const int* cols = tile_cols + t0.z;
        for (int e = tid; e < total; e += nth) {
            const int dcol = (int)__umulhi((unsigned)e, kinv);  // e / kt (exact for e < 2^16)
            const int c = e - dcol * kt;
            cp_async_16(&sV[e], V + (size_t)cols[dcol] * ldv + c);
        }
        asm volatile("cp.async.commit_group;" ::);
    }
    // 2. the tile's slice of the matrix stream (contiguous in CSR order)
    for (int e = tid; e < nnz; e += nth) {
        double v[VW];
        load_vals<VW>(vals, (size_t)(nz0 + e), v);
        if constexpr (DIAG) {
#pragma unroll
            for (int t = 0; t < VW; ++t) sM[e * VW + t] = v[t];
        } else {
            const double2 m = combine<VW, CA>(v, [&](int i) { return cp.c[i]; });
            *(double2*)(sM + 2 * e) = m;
        }
        sL[e] = lidx[nz0 + e];
    }
    if (tid <= nrows) sRp[tid] = rowptr[row0 + tid] - nz0;
    const int gc = tid % GC;
    const int r = tid / GC;
    double2 cd[DIAG ? CPT : 1][DIAG ? (CA ? VW / 2 : VW) : 1];
    if constexpr (DIAG) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
#pragma unroll
            for (int i = 0; i < (CA ? VW / 2 : VW); ++i) cd[j][i] = (c < kt) ? cdiag[i + p * c] : make_double2(0.0, 0.0);
        }
    }
    double2 acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = make_double2(0.0, 0.0);
    if constexpr (BULK) {
        unsigned done = 0, spins = 0;
        while (!done) {
            if (++spins > (1u << 26)) asm volatile("trap;");  // a lost transaction must fail the launch, never hang the device
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done)
                         : "r"(bar_s), "r"(0)
                         : "memory");
        }
    } else {
        asm volatile("cp.async.wait_group 0;" ::);
    }
    __syncthreads();
    int start = 0, end = 0;
    if (r < nrows) {
        start = sRp[r];
        end = sRp[r + 1];
    }
    // 3. row products
#pragma unroll 2
    for (int idx = start; idx < end; ++idx) {
        const double2* xr = sV + (int)sL[idx] * kt + gc;
        if constexpr (!DIAG) {
            const double2 m = *(const double2*)(sM + 2 * idx);
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (gc + j * GC < kt) cfma(acc[j], m, xr[j * GC]);
        } else {
            double v[VW];
#pragma unroll
            for (int t = 0; t < VW; ++t) v[t] = sM[idx * VW + t];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                if (gc + j * GC < kt) {
                    const double2 m = combine<VW, CA>(v, [&](int i) { return cd[j][i]; });
                    cfma(acc[j], m, xr[j * GC]);
                }
            }
        }
    }
    if (r < nrows) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
            if (c < kt) Z[(size_t)(row0 + r) * ldz + c] = acc[j];
        }
    }
}


// ---------------------------------------------------------------------------------------------
// K1b (round 2, the default for 5..32 dense columns): the tiled product with EVERY operand of a tile brought in by the TMA
// unit, so that the LSU / shared-memory pipe -- the measured limiter of spmm_tiled_kernel (profiles/r1_ncu_spmm_tiled.txt:
// l1tex 72 %, DRAM 36 %) -- only serves the row products:
//   * the tile's distinct columns are kept as RUNS of consecutive columns (host, spmf_build_tiles); V is row-major, so a run
//     of c consecutive rows of V is one contiguous piece of c*kt*16 bytes = ONE cp.async.bulk.  Neighbouring rows of a
//     discretisation share most of their columns: the C4 stencil has 5 runs of ~36 rows per 32-row tile instead of 180
//     separate row copies (a pattern without locality degenerates to one copy per row, still correct);
//   * the tile's slice of the interleaved values (contiguous in CSR order) and of the 16-bit tile-local column indices are
//     two more bulk copies (start addresses rounded down to 16 bytes, the slack skipped when reading);
//   * all copies complete on one mbarrier (transaction bytes); nothing is held in registers while the data is in flight, and
//     the other resident CTAs of the SM (4-7 of them) are in their product phase meanwhile -- the resident CTAs ARE the
//     pipeline stages.
// After the wait the raw values are combined once per nonzero (SCALAR) and the row products run as in spmm_tiled_kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(unsigned dst_s, const void* src, unsigned bytes, unsigned bar_s) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(src),
                 "r"(bytes), "r"(bar_s)
                 : "memory");
}

struct TmaSmem {  // byte offsets inside the dynamic shared memory of spmm_tma_kernel (host and device agree through this)
    unsigned sV, sA, sM, sL, sRp, bar, total;
};
__host__ __device__ inline TmaSmem tma_smem_layout(int max_cols, int max_nnz, int kt, int vw, bool diag, int rows) {
    TmaSmem L;
    unsigned o = 0;
    L.sV = o;
    o += (unsigned)max_cols * kt * 16;
    L.sA = o;
    o += ((unsigned)max_nnz * vw * 8 + 16 + 15) & ~15u;  // + slack for the rounded-down start
    L.sM = o;  // (unused since the combine pass was dropped)
    L.sL = o;
    o += ((unsigned)max_nnz * 2 + 16 + 15) & ~15u;
    L.sRp = o;
    o += ((unsigned)(rows + 1) * 4 + 15) & ~15u;
    L.bar = o;
    o += 16;
    L.total = o;
    return L;
}

__device__ int g_spmm_dbg = 0;  // timing experiments (NEPB_SPMM_DBG; wrong results): 1 no third V piece, 2 term values read once per trip, 4 no V loads
template <int VW, bool CA, bool DIAG, int CPT, int GC>
__global__ void __launch_bounds__(256) spmm_tma_kernel(int kt, int ldv, int ldz, int max_cols, int max_nnz, int tile_rows,
                                                       const int4* __restrict__ tiles, const int2* __restrict__ runs,
                                                       const int* __restrict__ tile_cols, const int* __restrict__ rowptr,
                                                       const uint16_t* __restrict__ lidx,
                                                       const double* __restrict__ vals, const double2* __restrict__ V,
                                                       double2* __restrict__ Z, const CoefP cp, const double2* __restrict__ cdiag, int p) {
    // GC lanes per row, lane gc owns the dense columns gc, gc + GC, ..  Every shared-memory wavefront DELIVERS at most 128
    // bytes to registers, so a 16-byte broadcast to the GC lanes of a row costs GC/8 of a wavefront per nonzero on top of the
    // k/8 wavefronts of the V row itself: GC = 4 halves the coefficient / index traffic of GC = 8 (profiles/r2_ncu_spmm_tma*.txt).
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TmaSmem L = tma_smem_layout(max_cols, max_nnz, kt, VW, DIAG, tile_rows);
    double2* sV = (double2*)(smem_raw + L.sV);
    int* sRp = (int*)(smem_raw + L.sRp);
    const int4 t0 = tiles[2 * blockIdx.x], t1 = tiles[2 * blockIdx.x + 1];
    const int row0 = t0.x, nrows = t0.y, ncols = t0.w, nz0 = t1.x, nnz = t1.y, run0 = t1.z, nruns = t1.w;
    const int tid = threadIdx.x, nth = blockDim.x;
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(smem_raw + L.bar);
    // slices of the matrix stream, start rounded down to 16 bytes
    const size_t a_byte = (size_t)nz0 * VW * 8, l_byte = (size_t)nz0 * 2;
    const unsigned a_skip = (unsigned)(a_byte & 15), l_skip = (unsigned)(l_byte & 15);
    const unsigned a_len = ((unsigned)nnz * VW * 8 + a_skip + 15) & ~15u, l_len = ((unsigned)nnz * 2 + l_skip + 15) & ~15u;
    const double* sA = (const double*)(smem_raw + L.sA + a_skip);
    const uint16_t* sL = (const uint16_t*)(smem_raw + L.sL + l_skip);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned total = (unsigned)ncols * kt * 16 + a_len + l_len;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(total) : "memory");
        bulk_g2s((unsigned)__cvta_generic_to_shared(smem_raw + L.sA), (const unsigned char*)vals + (a_byte - a_skip), a_len, bar_s);
        bulk_g2s((unsigned)__cvta_generic_to_shared(smem_raw + L.sL), (const unsigned char*)lidx + (l_byte - l_skip), l_len, bar_s);
    }
    if (ldv == kt) {
        for (int r = tid; r < nruns; r += nth) {  // one bulk copy per run of consecutive V rows
            const int2 rn = runs[run0 + r];
            const int dst_row = rn.y & 0xffff, len = rn.y >> 16;
            bulk_g2s((unsigned)__cvta_generic_to_shared(sV + (size_t)dst_row * kt), V + (size_t)rn.x * ldv, (unsigned)len * kt * 16, bar_s);
        }
    } else {  // a column window of a wider block: rows are not adjacent in memory, one copy per row
        const int* cols = tile_cols + t0.z;
        for (int dcol = tid; dcol < ncols; dcol += nth)
            bulk_g2s((unsigned)__cvta_generic_to_shared(sV + (size_t)dcol * kt), V + (size_t)cols[dcol] * ldv, (unsigned)kt * 16, bar_s);
    }
    if (tid <= nrows) sRp[tid] = rowptr[row0 + tid] - nz0;
    const int gc = tid % GC;
    const int r = tid / GC;
    double2 cd[DIAG ? CPT : 1][DIAG ? (CA ? VW / 2 : VW) : 1];
    if constexpr (DIAG) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
#pragma unroll
            for (int i = 0; i < (CA ? VW / 2 : VW); ++i) cd[j][i] = (c < kt) ? cdiag[i + p * c] : make_double2(0.0, 0.0);
        }
    }
    {
        unsigned done = 0, spins = 0;
        while (!done) {
            if (++spins > (1u << 26)) asm volatile("trap;");  // a lost transaction must fail the launch, never hang the device
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done)
                         : "r"(bar_s), "r"(0)
                         : "memory");
        }
    }
    __syncthreads();  // the row pointers
    int start = 0, end = 0;
    if (r < nrows) {
        start = sRp[r];
        end = sRp[r + 1];
    }
    // Row products straight from the raw term values (two 16-byte broadcast reads per nonzero for p = 4 real terms; no separate
    // combine pass: its strided shared-memory reads cost a full wavefront per nonzero, profiles/r2_ncu_spmm_tma.txt).
    // Full trips of UN nonzeros run without predicates (every load first, then the arithmetic, two accumulator sets so that
    // consecutive nonzeros do not form one dependent DFMA chain); the row's remainder is handled one nonzero at a time.
    // Lanes whose column lies beyond kt read whatever follows in shared memory and never store it.
    double2 acc[2][CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[0][j] = acc[1][j] = make_double2(0.0, 0.0);
    auto load_vals_s = [&](int idx, double (&v)[VW]) {
        const double* vp = sA + (size_t)idx * VW;
        if constexpr (VW % 2 == 0) {
#pragma unroll
            for (int t = 0; t < VW; t += 2) {
                double2 w;
                if (a_skip == 0) w = *(const double2*)(vp + t);  // uniform branch: 16-byte aligned slice
                else w = make_double2(vp[t], vp[t + 1]);
                v[t] = w.x;
                v[t + 1] = w.y;
            }
        } else {
#pragma unroll
            for (int t = 0; t < VW; ++t) v[t] = vp[t];
        }
    };
    auto load_x = [&](int li, double2 (&x)[CPT], int ncp) {
        const double2* xr = sV + li * kt + gc;
        if constexpr (GC == 8) {
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (j < ncp) x[j] = xr[j * GC];
        } else {
            // GC = 4: a quarter-warp (one LDS.128 wavefront) serves two rows, 64 bytes each; the two pieces must fall into
            // different halves of the 32 banks.  The half of piece j is (li*kt/4 + j) mod 2: neighbouring pieces are
            // read in swapped order when (row parity + li*kt/4) is odd, and swapped back in registers.
            const int sw = (r + ((li * kt) >> 2)) & 1;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int jj = ((j | 1) < CPT) ? (j ^ sw) : j;  // pairs (0,1), (2,3), ..; an odd last piece stays
                x[j] = (gc + jj * GC < kt) ? xr[jj * GC] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int j = 0; j + 1 < CPT; j += 2) {
                const double2 xa = x[j], xb = x[j + 1];
                x[j] = sw ? xb : xa;
                x[j + 1] = sw ? xa : xb;
            }
        }
    };
    auto fma_one = [&](const double (&v)[VW], const double2 (&x)[CPT], double2 (&ac)[CPT], int ncp) {
        if constexpr (!DIAG) {
            const double2 m = combine<VW, CA>(v, [&](int i) { return cp.c[i]; });
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (j < ncp) cfma(ac[j], m, x[j]);
        } else {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                if (j < ncp) {
                    const double2 m = combine<VW, CA>(v, [&](int i) { return cd[j][i]; });
                    cfma(ac[j], m, x[j]);
                }
            }
        }
    };
    // PAIR (17..20 columns, 8 lanes per row): the third 128-byte piece of a V row holds only columns 16..19, i.e. lanes 0-3.
    // Two consecutive nonzeros share ONE third load: lanes 0-3 take columns 16..19 of the first, lanes 4-7 those of the second
    // (separate accumulator, folded into lanes 0-3 at the end): 2.5 instead of 3 wavefronts per nonzero for the V rows.
    const bool pair = GC == 8 && CPT == 3 && !DIAG && kt <= 20;
    const int dbg = g_spmm_dbg;
    double2 acc3 = make_double2(0.0, 0.0);
    constexpr int UN = CPT >= 3 ? 2 : 4;
    int base = start;
    if (pair) {
#pragma unroll 2
        for (; base + 2 <= end; base += 2) {
            const int la = (int)sL[base], lb = (int)sL[base + 1];
            double va[VW], vb[VW];
            double2 xa[CPT], xb[CPT];
            load_vals_s(base, va);
            if (!(dbg & 2)) load_vals_s(base + 1, vb);
            else
                for (int t = 0; t < VW; ++t) vb[t] = va[t];
            if (!(dbg & 4)) {
                load_x(la, xa, 2);
                load_x(lb, xb, 2);
            } else {
                xa[0] = xa[1] = xb[0] = xb[1] = make_double2(1.0, (double)la + lb);
            }
            const double2 x3 = (dbg & 1) ? make_double2(1.0, 2.0) : sV[((gc < 4) ? la : lb) * kt + 16 + (gc & 3)];
            const double2 ma = combine<VW, CA>(va, [&](int i) { return cp.c[i]; });
            const double2 mb = combine<VW, CA>(vb, [&](int i) { return cp.c[i]; });
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                cfma(acc[0][j], ma, xa[j]);
                cfma(acc[1][j], mb, xb[j]);
            }
            cfma(acc3, (gc < 4) ? ma : mb, x3);
        }
    } else {
        for (; base + UN <= end; base += UN) {
            int li[UN];
            double v[UN][VW];
            double2 x[UN][CPT];
#pragma unroll
            for (int u = 0; u < UN; ++u) li[u] = (int)sL[base + u];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                load_vals_s(base + u, v[u]);
                load_x(li[u], x[u], CPT);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) fma_one(v[u], x[u], acc[u & 1], CPT);
        }
    }
    for (; base < end; ++base) {
        double v[VW];
        double2 x[CPT];
        load_vals_s(base, v);
        load_x((int)sL[base], x, CPT);
        fma_one(v, x, acc[0], CPT);
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        acc[0][j].x += acc[1][j].x;
        acc[0][j].y += acc[1][j].y;
    }
    if (pair) {  // lanes 4-7 hand their share of columns 16..19 to lanes 0-3 of the same row
        const double tx = __shfl_down_sync(0xffffffffu, acc3.x, 4), ty = __shfl_down_sync(0xffffffffu, acc3.y, 4);
        if (gc < 4) {
            acc[0][2].x += acc3.x + tx;
            acc[0][2].y += acc3.y + ty;
        }
    }
    if (r < nrows) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
            if (c < kt) Z[(size_t)(row0 + r) * ldz + c] = acc[0][j];
        }
    }
}

// runtime-p fallback (any p <= MAXP, real or complex values); same row ownership with GC=4, GN=2
template <bool DIAG>
__global__ void __launch_bounds__(256) spmm_fused_generic_kernel(int n, int kt, int ldv, int ldz, const int* __restrict__ rowptr,
                                                                 const int* __restrict__ colind, const double* __restrict__ vals,
                                                                 const double2* __restrict__ V, double2* __restrict__ Z,
                                                                 const CoefP cp, const double2* __restrict__ cdiag, int p, int ca) {
    constexpr int GC = 4, GN = 2, G = 8, ROWS = 256 / G, CPT = 8;
    const int tid = threadIdx.x;
    const int g = tid % G, gc = g % GC, gn = g / GC;
    const int row = blockIdx.x * ROWS + tid / G;
    const int vw = ca ? 2 * p : p;
    double2 acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = make_double2(0.0, 0.0);
    int start = 0, end = 0;
    if (row < n) {
        start = rowptr[row];
        end = rowptr[row + 1];
    }
    for (int idx = start + gn; idx < end; idx += GN) {
        const int col = colind[idx];
        const double* v = vals + (size_t)idx * vw;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
            if (c < kt) {
                double2 m = make_double2(0.0, 0.0);
                for (int i = 0; i < p; ++i) {
                    const double2 ci = DIAG ? cdiag[i + p * c] : cp.c[i];
                    const double2 a = ca ? make_double2(v[2 * i], v[2 * i + 1]) : make_double2(v[i], 0.0);
                    cfma(m, ci, a);
                }
                cfma(acc[j], m, __ldg(V + (size_t)col * ldv + c));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, GC);
        acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, GC);
    }
    if (row < n && gn == 0) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
            if (c < kt) Z[(size_t)row * ldz + c] = acc[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: GENERAL mode, second stage.  X = V * [C_1 .. C_p] was formed by the panel kernel below with
// layout X[row][c][i] (the p terms of output column c contiguous), then
//   Z[row, c] = sum_nz sum_i val_i(nz) * X[col(nz), c, i].
// ---------------------------------------------------------------------------------------------
template <int VW, bool CA, int GC, int GN, int U>
__global__ void __launch_bounds__(256) spmm_stacked_kernel(int n, int q, const int* __restrict__ rowptr,
                                                           const int* __restrict__ colind, const double* __restrict__ vals,
                                                           const double2* __restrict__ X, double2* __restrict__ Z, int ldz) {
    constexpr int P = CA ? VW / 2 : VW;
    constexpr int G = GC * GN;
    constexpr int ROWS = 256 / G;
    const int tid = threadIdx.x;
    const int g = tid % G, gc = g % GC, gn = g / GC;
    const int row = blockIdx.x * ROWS + tid / G;
    int start = 0, end = 0;
    if (row < n) {
        start = rowptr[row];
        end = rowptr[row + 1];
    }
    for (int c0 = 0; c0 < q; c0 += GC) {  // uniform trip count for the whole grid
        const int c = c0 + gc;
        double2 acc = make_double2(0.0, 0.0);
        for (int base = start; base < end; base += GN * U) {
            int col[U];
            double v[U][VW];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * GN + gn;
                ok[u] = idx < end && c < q;
                col[u] = 0;
                if (ok[u]) {
                    col[u] = ld_stream_s32(colind + idx);
                    load_vals<VW>(vals, (size_t)idx, v[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) {
                    const double2* xp = X + ((size_t)col[u] * q + c) * P;
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const double2 x = __ldg(xp + i);
                        if constexpr (CA) {
                            cfma(acc, make_double2(v[u][2 * i], v[u][2 * i + 1]), x);
                        } else {
                            acc.x = fma(v[u][i], x.x, acc.x);
                            acc.y = fma(v[u][i], x.y, acc.y);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int off = GC; off < G; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (row < n && gn == 0 && c < q) Z[(size_t)row * ldz + c] = acc;
    }
}

__global__ void __launch_bounds__(256) spmm_stacked_generic_kernel(int n, int q, int p, int ca, const int* __restrict__ rowptr,
                                                                   const int* __restrict__ colind, const double* __restrict__ vals,
                                                                   const double2* __restrict__ X, double2* __restrict__ Z, int ldz) {
    constexpr int G = 8, ROWS = 256 / G;
    const int tid = threadIdx.x;
    const int gn = tid % G;
    const int row = blockIdx.x * ROWS + tid / G;
    const int vw = ca ? 2 * p : p;
    int start = 0, end = 0;
    if (row < n) {
        start = rowptr[row];
        end = rowptr[row + 1];
    }
    for (int c = 0; c < q; ++c) {
        double2 acc = make_double2(0.0, 0.0);
        for (int idx = start + gn; idx < end; idx += G) {
            const int col = colind[idx];
            const double* v = vals + (size_t)idx * vw;
            const double2* xp = X + ((size_t)col * q + c) * p;
            for (int i = 0; i < p; ++i) {
                const double2 a = ca ? make_double2(v[2 * i], v[2 * i + 1]) : make_double2(v[i], 0.0);
                cfma(acc, a, __ldg(xp + i));
            }
        }
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (row < n && gn == 0) Z[(size_t)row * ldz + c] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// K0: tall-skinny panel product X(n x w) = V(n x k) * Cs(k x w), all row-major complex.
// 32 rows x 32 k-chunk tile of V staged in shared memory with coalesced loads; bandwidth-bound on V.
// ---------------------------------------------------------------------------------------------
constexpr int PANEL_R = 32, PANEL_K = 32, PANEL_W = 8;  // block computes 32 rows x (8 outputs per pass)
__global__ void __launch_bounds__(256) panel_gemm_kernel(int n, int k, int w, const double2* __restrict__ V, int ldv,
                                                         const double2* __restrict__ Cs, double2* __restrict__ X) {
    __shared__ double2 sV[PANEL_R][PANEL_K + 1];
    __shared__ double2 sC[PANEL_K][PANEL_W];
    const int tid = threadIdx.x;
    const int r = tid / PANEL_W;  // 0..31
    const int cw = tid % PANEL_W;
    const int row0 = blockIdx.x * PANEL_R;
    for (int w0 = 0; w0 < w; w0 += PANEL_W) {
        double2 acc = make_double2(0.0, 0.0);
        for (int k0 = 0; k0 < k; k0 += PANEL_K) {
            __syncthreads();
            for (int t = tid; t < PANEL_R * PANEL_K; t += 256) {
                const int rr = t / PANEL_K, kk = t % PANEL_K;
                double2 val = make_double2(0.0, 0.0);
                if (row0 + rr < n && k0 + kk < k) val = V[(size_t)(row0 + rr) * ldv + k0 + kk];
                sV[rr][kk] = val;
            }
            {
                const int kk = tid / PANEL_W, ww = tid % PANEL_W;
                double2 val = make_double2(0.0, 0.0);
                if (k0 + kk < k && w0 + ww < w) val = Cs[(size_t)(k0 + kk) * w + w0 + ww];
                sC[kk][ww] = val;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < PANEL_K; ++kk) cfma(acc, sV[r][kk], sC[kk][cw]);
        }
        if (row0 + r < n && w0 + cw < w) X[(size_t)(row0 + r) * w + w0 + cw] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// GENERAL mode, round 2.  Stage 1: X(n x w) = V(n x k) * Cs(k x w), w = p*q small.  Bandwidth bound on V: 8 lanes per row
// (a warp instruction reads four rows with 128 contiguous bytes each), lane g owns k-indices g, g + 8, ..; the WT outputs of a
// pass are reduced over the 8 lanes with 3 shuffle steps and written as one contiguous piece per row.  Cs sits in shared
// memory (k*w*16 bytes).  Round 1's 32 x 32 shared-memory tile kernel took 219 us where 60 us is the HBM time (C4, k = 20).
// ---------------------------------------------------------------------------------------------
template <int WT>
__global__ void __launch_bounds__(256) panel_rows_kernel(int n, int k, int w, const double2* __restrict__ V, int ldv,
                                                         const double2* __restrict__ Cs, double2* __restrict__ X) {
    // 4 lanes per row (lane g owns the k-indices g, g + 4, ..), 4 rows per lane: a coefficient Cs[kk][c] read from shared memory
    // serves 4 rows, and the reduction over the lanes of a row needs 2 shuffle steps.  (First version: 8 lanes per row, one row per
    // lane -- the coefficient reads alone kept the shared-memory pipe 83-90 % busy, DRAM at 20 %.)  Pitch w + 1: the 8 lanes of
    // a 128-bit shared-memory phase read 4 different rows of Cs, conflict-free with the odd pitch.
    extern __shared__ double2 sCs[];  // [k][w + 1]
    const int pw = w + 1;
    for (int idx = threadIdx.x; idx < k * w; idx += 256) sCs[(idx / w) * pw + idx % w] = Cs[idx];
    __syncthreads();
    constexpr int RPL = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane & 3, rg = lane >> 2;
    const int row_base = blockIdx.x * (8 * 8 * RPL) + warp * (8 * RPL) + rg;  // rows row_base + 8 u, u < RPL
    for (int w0 = 0; w0 < w; w0 += WT) {
        double2 acc[RPL][WT];
#pragma unroll
        for (int u = 0; u < RPL; ++u)
#pragma unroll
            for (int c = 0; c < WT; ++c) acc[u][c] = make_double2(0.0, 0.0);
        for (int kk = g; kk < k; kk += 4) {
            double2 v[RPL], cc[WT];
#pragma unroll
            for (int u = 0; u < RPL; ++u) {
                const int r = row_base + 8 * u;
                v[u] = r < n ? V[(size_t)r * ldv + kk] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int c = 0; c < WT; ++c) cc[c] = (w0 + c < w) ? sCs[kk * pw + w0 + c] : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < RPL; ++u)
#pragma unroll
                for (int c = 0; c < WT; ++c) cfma(acc[u][c], v[u], cc[c]);
        }
#pragma unroll
        for (int u = 0; u < RPL; ++u)
#pragma unroll
            for (int c = 0; c < WT; ++c) {
                acc[u][c].x += __shfl_xor_sync(0xffffffffu, acc[u][c].x, 1);
                acc[u][c].y += __shfl_xor_sync(0xffffffffu, acc[u][c].y, 1);
                acc[u][c].x += __shfl_xor_sync(0xffffffffu, acc[u][c].x, 2);
                acc[u][c].y += __shfl_xor_sync(0xffffffffu, acc[u][c].y, 2);
            }
#pragma unroll
        for (int u = 0; u < RPL; ++u) {
            const int r = row_base + 8 * u;
            if (r < n) {
#pragma unroll
                for (int c = 0; c < WT; ++c)
                    if ((c & 3) == g && w0 + c < w) X[(size_t)r * w + w0 + c] = acc[u][c];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// GENERAL mode, stage 2 (round 2): Z[row, c] = sum_nz sum_i a_i(nz) X[col(nz)][c][i] with the rows of X (p*q complex values,
// contiguous) staged per 16-row tile by TMA bulk copies of column runs, like spmm_tma_kernel; real term values, p <= 4.
// Lane layout per row: 4 lanes = the 4 terms (lane i multiplies with a_i: the 32 bytes of term values of a nonzero are read as
// four different 8-byte pieces, nothing is broadcast), times QS column slots (1 for q = 1, else 2).  The term sums are
// reduced over the 4 lanes with 2 shuffle steps at the end of the row.
// ---------------------------------------------------------------------------------------------
template <int QS, int CPT>
__global__ void __launch_bounds__(128) spmm_stacked_tma_kernel(int q, int p, int ldz, int max_cols, int max_nnz, int tile_rows,
                                                               const int4* __restrict__ tiles, const int2* __restrict__ runs,
                                                               const int* __restrict__ rowptr, const uint16_t* __restrict__ lidx,
                                                               const double* __restrict__ vals, const double2* __restrict__ X,
                                                               double2* __restrict__ Z) {
    constexpr int LPR = 4 * QS;  // lanes per row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = p * q;  // complex values per row of X
    const TmaSmem L = tma_smem_layout(max_cols, max_nnz, P, 4, true, tile_rows);
    const double2* sX = (const double2*)(smem_raw + L.sV);
    int* sRp = (int*)(smem_raw + L.sRp);
    const int4 t0 = tiles[2 * blockIdx.x], t1 = tiles[2 * blockIdx.x + 1];
    const int row0 = t0.x, nrows = t0.y, ncols = t0.w, nz0 = t1.x, nnz = t1.y, run0 = t1.z, nruns = t1.w;
    const int tid = threadIdx.x, nth = blockDim.x;
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(smem_raw + L.bar);
    const size_t a_byte = (size_t)nz0 * p * 8, l_byte = (size_t)nz0 * 2;  // p doubles per nonzero (real values)
    const unsigned a_skip = (unsigned)(a_byte & 15), l_skip = (unsigned)(l_byte & 15);
    const unsigned a_len = ((unsigned)nnz * p * 8 + a_skip + 15) & ~15u, l_len = ((unsigned)nnz * 2 + l_skip + 15) & ~15u;
    const double* sA = (const double*)(smem_raw + L.sA + a_skip);
    const uint16_t* sL = (const uint16_t*)(smem_raw + L.sL + l_skip);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned total = (unsigned)ncols * P * 16 + a_len + l_len;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(total) : "memory");
        bulk_g2s((unsigned)__cvta_generic_to_shared(smem_raw + L.sA), (const unsigned char*)vals + (a_byte - a_skip), a_len, bar_s);
        bulk_g2s((unsigned)__cvta_generic_to_shared(smem_raw + L.sL), (const unsigned char*)lidx + (l_byte - l_skip), l_len, bar_s);
    }
    for (int r = tid; r < nruns; r += nth) {
        const int2 rn = runs[run0 + r];
        const int dst_row = rn.y & 0xffff, len = rn.y >> 16;
        bulk_g2s((unsigned)__cvta_generic_to_shared(sX + (size_t)dst_row * P), X + (size_t)rn.x * P, (unsigned)len * P * 16, bar_s);
    }
    if (tid <= nrows) sRp[tid] = rowptr[row0 + tid] - nz0;
    {
        unsigned done = 0, spins = 0;
        while (!done) {
            if (++spins > (1u << 26)) asm volatile("trap;");
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done)
                         : "r"(bar_s), "r"(0)
                         : "memory");
        }
    }
    __syncthreads();
    const int i = tid & 3;                   // term
    const int cs = (tid % LPR) >> 2;         // column slot
    const int r = tid / LPR;                 // row within the tile
    int start = 0, end = 0;
    if (r < nrows) {
        start = sRp[r];
        end = sRp[r + 1];
    }
    double2 acc[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[j] = make_double2(0.0, 0.0);
    const bool term_ok = i < p;
#pragma unroll 2
    for (int idx = start; idx < end; ++idx) {
        const double a = term_ok ? sA[(size_t)idx * p + i] : 0.0;   // p doubles per nonzero (vw = p, real values)
        const double2* xr = sX + (size_t)sL[idx] * P + (term_ok ? i : 0);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = cs + j * QS;
            if (c < q) {
                const double2 x = xr[c * p];
                acc[j].x = fma(a, x.x, acc[j].x);
                acc[j].y = fma(a, x.y, acc[j].y);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, 1);
        acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, 1);
        acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, 2);
        acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, 2);
    }
    if (r < nrows && i == 0) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = cs + j * QS;
            if (c < q) Z[(size_t)(row0 + r) * ldz + c] = acc[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Mder values: out[e] (CSC order) = sum_i c_i * vals[csr_of_csc[e]][i]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mder_kernel(int64_t nnz, int p, int ca, const int* __restrict__ csr_of_csc,
                                                   const double* __restrict__ vals, const CoefP cp, double2* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int vw = ca ? 2 * p : p;
    const double* v = vals + (size_t)csr_of_csc[e] * vw;
    double2 m = make_double2(0.0, 0.0);
    for (int i = 0; i < p; ++i) {
        const double2 a = ca ? make_double2(v[2 * i], v[2 * i + 1]) : make_double2(v[i], 0.0);
        cfma(m, cp.c[i], a);
    }
    out[e] = m;
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
struct TileCfg {
    int gc, gn, cpt, u;
};

static bool env_cfg(TileCfg& c) {
    const char* s = getenv("NEPB_SPMM_CFG");
    if (!s) return false;
    return sscanf(s, "%d,%d,%d,%d", &c.gc, &c.gn, &c.cpt, &c.u) == 4;
}

static TileCfg default_cfg(int kt) {
    if (kt <= 1) return {1, 8, 1, 4};
    if (kt <= 2) return {2, 4, 1, 4};
    if (kt <= 4) return {4, 2, 1, 4};
    // multi-column shapes: measured on config C4 (profiles/r1_spmm_tile_sweep.txt): shallow unrolling wins, the kernel is
    // occupancy / gather-latency bound rather than HBM bound once k*16 B rows are gathered through L1/L2
    if (kt <= 8) return {4, 4, 2, 2};
    if (kt <= 16) return {4, 2, 4, 1};
    if (kt <= 20) return {4, 2, 5, 1};
    return {8, 1, 4, 4};
}

#define NEPB_TILE_LIST(X) \
    X(1, 8, 1, 4) X(2, 4, 1, 4) X(4, 2, 1, 4) X(4, 4, 2, 2) X(4, 2, 4, 1) X(4, 2, 5, 1) X(8, 1, 4, 4)
// extra shapes kept for on-device tuning of the benchmark case (p=4 real, SCALAR)
#define NEPB_TUNE_LIST(X)                                                                              \
    X(1, 4, 1, 6) X(1, 8, 1, 3) X(1, 16, 1, 2) X(1, 32, 1, 1) X(1, 4, 1, 8) X(8, 1, 1, 4) X(8, 2, 1, 4) \
    X(8, 1, 1, 8) X(4, 1, 5, 4) X(4, 4, 5, 1) X(4, 2, 5, 4) X(8, 1, 3, 4) X(8, 2, 3, 2) X(2, 4, 10, 2) X(4, 2, 2, 4) X(4, 2, 4, 2) X(4, 2, 5, 2) \
    X(2, 8, 4, 2) X(4, 4, 2, 4) X(4, 8, 2, 1) X(8, 4, 1, 2) X(4, 8, 2, 2) X(2, 16, 4, 1) X(4, 4, 2, 1) X(8, 4, 1, 1) \
    X(4, 4, 5, 2) X(4, 8, 5, 1) X(2, 8, 10, 1) X(8, 4, 3, 1) X(8, 2, 3, 1) X(4, 8, 5, 2)

template <int VW, bool CA, bool DIAG>
static int launch_fused_vw(const nepb_spmf* h, TileCfg cfg, int kt, int ldv, int ldz, const double2* V, double2* Z,
                           const CoefP& cp, const double2* cdiag, bool allow_tune) {
    const int n = (int)h->n;
#define NEPB_TRY(GC_, GN_, CPT_, U_)                                                                           \
    if (cfg.gc == GC_ && cfg.gn == GN_ && cfg.cpt == CPT_ && cfg.u == U_) {                                     \
        const int rows = 256 / (GC_ * GN_);                                                                     \
        NEPB_LAUNCH((spmm_fused_kernel<VW, CA, DIAG, GC_, GN_, CPT_, U_>), (n + rows - 1) / rows, 256, 0, n, kt, \
                    ldv, ldz, h->d_rowptr.p, h->d_colind.p, h->d_vals.p, V, Z, cp, cdiag, h->p);                \
        return 1;                                                                                               \
    }
    NEPB_TILE_LIST(NEPB_TRY)
    if constexpr (VW == 4 && !CA && !DIAG) {
        if (allow_tune) {
            NEPB_TUNE_LIST(NEPB_TRY)
        }
    }
#undef NEPB_TRY
    return 0;
}

// Cut the rows into tiles for spmm_tiled_kernel (host, once per operator and tile height, O(nnz)): greedy over consecutive
// rows inside fixed chunks of 8192 rows (one chunk per OpenMP task); a tile closes at `tile_rows` rows or when the next row
// would push its distinct columns beyond 6 * tile_rows.  Pure integer work, deterministic.
static int spmf_build_tiles(const nepb_spmf* h, int which) {
    nepb_spmf::TileSet& T = h->tiling[which];
    if (T.state) return T.state;
    const int tile_rows = which == 0 ? 32 : 16;
    const int tile_cols_max = which == 0 ? 6 * tile_rows : 7 * tile_rows;  // 16 rows of a 5-line stencil touch 100 columns
    const int64_t n = h->n;
    const int32_t* rp = h->h_rowptr;
    const int32_t* ci = h->h_colind;
    constexpr int64_t CHUNK = 8192;
    const int64_t nchunks = (n + CHUNK - 1) / CHUNK;
    std::vector<std::vector<int4>> ctiles(nchunks);
    std::vector<std::vector<int32_t>> ccols(nchunks);
    std::vector<std::vector<int2>> cruns(nchunks);  // runs of consecutive columns: (first column, position in the tile | length << 16)
    std::vector<uint16_t> lidx((size_t)h->nnz);
    int bad = 0;
#pragma omp parallel
    {
        std::vector<int32_t> stamp((size_t)n, -1), local((size_t)n, 0), cur;
#pragma omp for schedule(dynamic, 1) reduction(| : bad)
        for (int64_t ch = 0; ch < nchunks; ++ch) {
            const int64_t r_end = std::min(n, (ch + 1) * CHUNK);
            int64_t r = ch * CHUNK;
            while (r < r_end) {
                cur.clear();
                const int32_t tile_id = (int32_t)r;  // stamps are first rows of tiles: unique
                int64_t rr = r;
                while (rr < r_end && rr - r < tile_rows) {
                    int added = 0;
                    for (int32_t e = rp[rr]; e < rp[rr + 1]; ++e)
                        if (stamp[ci[e]] != tile_id) ++added;
                    if ((int)cur.size() + added > tile_cols_max) break;
                    for (int32_t e = rp[rr]; e < rp[rr + 1]; ++e)
                        if (stamp[ci[e]] != tile_id) {
                            stamp[ci[e]] = tile_id;
                            cur.push_back(ci[e]);
                        }
                    ++rr;
                }
                if (rr == r) {  // a single row with more distinct columns than a tile holds
                    bad |= 1;
                    rr = r + 1;
                    cur.clear();
                }
                std::sort(cur.begin(), cur.end());
                for (size_t t = 0; t < cur.size(); ++t) local[cur[t]] = (int32_t)t;
                const int run_first = (int)cruns[ch].size();
                for (size_t t = 0; t < cur.size();) {
                    size_t u = t + 1;
                    while (u < cur.size() && cur[u] == cur[u - 1] + 1) ++u;
                    cruns[ch].push_back(make_int2(cur[t], (int)t | ((int)(u - t) << 16)));
                    t = u;
                }
                ctiles[ch].push_back(make_int4((int)r, (int)(rr - r), (int)ccols[ch].size(), (int)cur.size()));
                ctiles[ch].push_back(make_int4(rp[r], rp[rr] - rp[r], run_first, (int)cruns[ch].size() - run_first));
                ccols[ch].insert(ccols[ch].end(), cur.begin(), cur.end());
                if (!(bad & 1))
                    for (int32_t e = rp[r]; e < rp[rr]; ++e) lidx[e] = (uint16_t)local[ci[e]];
                r = rr;
            }
        }
    }
    if (bad) {
        T.state = -1;
        return -1;
    }
    std::vector<int4> tiles;
    std::vector<int32_t> cols;
    std::vector<int2> runs;
    int maxc = 0, maxz = 0;
    for (int64_t ch = 0; ch < nchunks; ++ch) {
        const int off = (int)cols.size();
        const int roff = (int)runs.size();
        runs.insert(runs.end(), cruns[ch].begin(), cruns[ch].end());
        for (size_t t = 0; t < ctiles[ch].size(); t += 2) {
            int4 a = ctiles[ch][t];
            a.z += off;
            tiles.push_back(a);
            int4 b = ctiles[ch][t + 1];
            b.z += roff;
            tiles.push_back(b);
            maxc = std::max(maxc, a.w);
            maxz = std::max(maxz, ctiles[ch][t + 1].y);
        }
        cols.insert(cols.end(), ccols[ch].begin(), ccols[ch].end());
    }
    if (cols.size() >= (size_t)std::numeric_limits<int32_t>::max()) {
        T.state = -1;
        return -1;
    }
    cudaError_t e = T.tiles.alloc(std::max<size_t>(tiles.size(), 1));
    if (e == cudaSuccess && !tiles.empty()) e = cudaMemcpy(T.tiles.p, tiles.data(), sizeof(int4) * tiles.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = T.cols.alloc(std::max<size_t>(cols.size(), 1));
    if (e == cudaSuccess && !cols.empty()) e = cudaMemcpy(T.cols.p, cols.data(), sizeof(int32_t) * cols.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = T.lidx.alloc(lidx.size() + 16);  // + slack: bulk copies round the slice end up to 16 bytes
    if (e == cudaSuccess && !lidx.empty()) e = cudaMemcpy(T.lidx.p, lidx.data(), sizeof(uint16_t) * lidx.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = T.runs.alloc(std::max<size_t>(runs.size(), 1));
    if (e == cudaSuccess && !runs.empty()) e = cudaMemcpy(T.runs.p, runs.data(), sizeof(int2) * runs.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaGetLastError();
        T.tiles.release();
        T.cols.release();
        T.lidx.release();
        T.runs.release();
        T.state = -1;  // not enough memory for the tile index: the untiled kernels still work
        return -1;
    }
    T.ntiles = (int64_t)tiles.size() / 2;
    T.runs_total = (int64_t)runs.size();
    T.cols_total = (int64_t)cols.size();
    T.max_cols = maxc;
    T.max_nnz = maxz;
    T.state = 1;
    return 1;
}

static size_t tiled_smem_bytes(const nepb_spmf::TileSet& T, int kt, int mw) {
    return (size_t)T.max_cols * kt * 16 + (size_t)T.max_nnz * mw * 8 + (size_t)((T.max_nnz + 7) & ~7) * 2 + 36 * 4 + 16;
}

template <int VW, bool CA, bool DIAG>
static int launch_tiled_vw(const nepb_spmf* h, const nepb_spmf::TileSet& T, int threads, int kt, int ldv, int ldz, const double2* V,
                           double2* Z, const CoefP& cp, const double2* cdiag) {
    const unsigned kinv = (unsigned)((0x100000000ULL + (unsigned)kt - 1) / (unsigned)kt);
    const size_t smem = tiled_smem_bytes(T, kt, DIAG ? VW : 2);
    const bool bulk = getenv("NEPB_SPMM_BULK") && atoi(getenv("NEPB_SPMM_BULK")) != 0;  // read per call: tests toggle it
#define NEPB_TILED_B(CPT_, BULK_)                                                                                                     \
    do {                                                                                                                              \
        static size_t attr_done[16] = {0}; /* per device */                                                                           \
        int dev_ = 0;                                                                                                                 \
        cudaGetDevice(&dev_);                                                                                                         \
        if (smem > 48 * 1024 && smem > attr_done[dev_ & 15]) {                                                                        \
            NEPB_CUDA(cudaFuncSetAttribute(spmm_tiled_kernel<VW, CA, DIAG, CPT_, BULK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem));                                                                               \
            attr_done[dev_ & 15] = smem;                                                                                              \
        }                                                                                                                             \
        NEPB_LAUNCH((spmm_tiled_kernel<VW, CA, DIAG, CPT_, BULK_>), (unsigned)T.ntiles, threads, smem, kt, kinv, ldv, ldz, T.max_cols, \
                    T.max_nnz, T.tiles.p, T.cols.p, h->d_rowptr.p, T.lidx.p, h->d_vals.p, V, Z, cp, cdiag, h->p);                     \
    } while (0)
#define NEPB_TILED(CPT_)                      \
    do {                                      \
        if (bulk) NEPB_TILED_B(CPT_, true);   \
        else NEPB_TILED_B(CPT_, false);       \
    } while (0)
    if (kt <= 8) NEPB_TILED(1);
    else if (kt <= 16) NEPB_TILED(2);
    else if (kt <= 24) NEPB_TILED(3);
    else NEPB_TILED(4);
#undef NEPB_TILED
#undef NEPB_TILED_B
    return 1;
}

template <int VW, bool CA, bool DIAG>
static int launch_tma_vw(const nepb_spmf* h, const nepb_spmf::TileSet& T, int tile_rows, int gc, int kt, int ldv, int ldz, const double2* V,
                         double2* Z, const CoefP& cp, const double2* cdiag) {
    const TmaSmem L = tma_smem_layout(T.max_cols, T.max_nnz, kt, VW, DIAG, tile_rows);
#define NEPB_TMA(CPT_, GC_)                                                                                                             \
    do {                                                                                                                                \
        static size_t attr_done[16] = {0};                                                                                              \
        int dev = 0;                                                                                                                    \
        cudaGetDevice(&dev);                                                                                                            \
        if (L.total > 48 * 1024 && L.total > attr_done[dev & 15]) {                                                                     \
            NEPB_CUDA(cudaFuncSetAttribute(spmm_tma_kernel<VW, CA, DIAG, CPT_, GC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total)); \
            attr_done[dev & 15] = L.total;                                                                                              \
        }                                                                                                                               \
        NEPB_LAUNCH((spmm_tma_kernel<VW, CA, DIAG, CPT_, GC_>), (unsigned)T.ntiles, GC_ * tile_rows, L.total, kt, ldv, ldz, T.max_cols, \
                    T.max_nnz, tile_rows, T.tiles.p, T.runs.p, T.cols.p, h->d_rowptr.p, T.lidx.p, h->d_vals.p, V, Z, cp, cdiag, h->p);  \
    } while (0)
    if (gc == 4) {
        if (kt <= 4) NEPB_TMA(1, 4);
        else if (kt <= 8) NEPB_TMA(2, 4);
        else if (kt <= 12) NEPB_TMA(3, 4);
        else if (kt <= 16) NEPB_TMA(4, 4);
        else if (kt <= 20) NEPB_TMA(5, 4);
        else if (kt <= 24) NEPB_TMA(6, 4);
        else NEPB_TMA(8, 4);
    } else {
        if (kt <= 8) NEPB_TMA(1, 8);
        else if (kt <= 16) NEPB_TMA(2, 8);
        else if (kt <= 24) NEPB_TMA(3, 8);
        else NEPB_TMA(4, 8);
    }
#undef NEPB_TMA
    return 1;
}


// ---------------------------------------------------------------------------------------------
// Two-dimensional tiles (round 2).  Measured (profiles/r2_spmm_2d.txt): with the row products' shared-memory reads switched
// off the 16-row kernel above still takes 404 of 430 us at k = 20 -- the product is bound by the bytes that travel from L2
// into the SMs (~6 TB/s for every variant), and a strip of 16 consecutive rows of a five-line stencil stages 5.6 rows of V per
// matrix row.  A tile of S = 4 segments x R = 8 rows, the segments one dominant column offset ("grid line") apart, shares
// its lines: 8 lines x 12 columns for 32 rows = 3.0 rows of V per matrix row (1.9 -> 1.0 GB of staging at k = 20).
// The line length is found from the pattern (most frequent |column - row| >= 32); patterns without one keep the 1D tiles.
// SCALAR mode, real term values with an even count (16-byte aligned slices); everything else as in spmm_tma_kernel.
// ---------------------------------------------------------------------------------------------
template <int VW, int CPT, bool PRE>
__global__ void __launch_bounds__(256) spmm_tma2d_kernel(int kt, int ldv, int ldz, int max_cols, int max_nnz, int S, int R,
                                                         const int4* __restrict__ desc, const int2* __restrict__ runs,
                                                         const int* __restrict__ tile_cols, const int* __restrict__ rowptr,
                                                         const uint16_t* __restrict__ lidx, const double* __restrict__ vals,
                                                         const double2* __restrict__ V, double2* __restrict__ Z, const CoefP cp) {
    constexpr int GC = 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned oA = (unsigned)max_cols * kt * 16, oL = oA + (unsigned)max_nnz * VW * 8, oB = oL + (((unsigned)max_nnz + 7) & ~7u) * 2;
    double2* sV = (double2*)smem_raw;
    const double* sA = (const double*)(smem_raw + oA);
    const uint16_t* sL = (const uint16_t*)(smem_raw + oL);
    const int4* d = desc + (size_t)blockIdx.x * (2 + S);
    const int4 h0 = d[0], h1 = d[1];
    const int ncols = h0.y, run0 = h0.z, nruns = h0.w, nnz = h1.y, nseg = h1.z;
    const int tid = threadIdx.x, nth = blockDim.x;
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(smem_raw + oB);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::);
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned lbytes = (((unsigned)nnz + 7) & ~7u) * 2;
        const unsigned total = (unsigned)ncols * kt * 16 + (unsigned)nnz * VW * 8 + lbytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(total) : "memory");
        unsigned off = 0;
        for (int s = 0; s < nseg; ++s) {  // the segments' value slices, back to back
            const int4 sg = d[2 + s];
            if (sg.w > 0) bulk_g2s((unsigned)__cvta_generic_to_shared(smem_raw + oA) + off, vals + (size_t)sg.z * VW, (unsigned)sg.w * VW * 8, bar_s);
            off += (unsigned)sg.w * VW * 8;
        }
        if (lbytes) bulk_g2s((unsigned)__cvta_generic_to_shared(smem_raw + oL), lidx + h1.x, lbytes, bar_s);
    }
    if (ldv == kt) {
        for (int r = tid; r < nruns; r += nth) {
            const int2 rn = runs[run0 + r];
            const int dst_row = rn.y & 0xffff, len = rn.y >> 16;
            bulk_g2s((unsigned)__cvta_generic_to_shared(sV + (size_t)dst_row * kt), V + (size_t)rn.x * ldv, (unsigned)len * kt * 16, bar_s);
        }
    } else {
        const int* cols = tile_cols + h0.x;
        for (int dcol = tid; dcol < ncols; dcol += nth)
            bulk_g2s((unsigned)__cvta_generic_to_shared(sV + (size_t)dcol * kt), V + (size_t)cols[dcol] * ldv, (unsigned)kt * 16, bar_s);
    }
    const int gc = tid % GC, t = tid / GC;
    const int sgi = t / R, ri = t - sgi * R;
    int start = 0, end = 0, zrow = -1;
    if (sgi < nseg) {
        int off = 0;
        for (int s = 0; s < sgi; ++s) off += d[2 + s].w;
        const int4 sg = d[2 + sgi];
        if (ri < sg.y) {
            zrow = sg.x + ri;
            start = off + rowptr[zrow] - sg.z;
            end = off + rowptr[zrow + 1] - sg.z;
        }
    }
    {
        unsigned done = 0, spins = 0;
        while (!done) {
            if (++spins > (1u << 26)) asm volatile("trap;");
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done)
                         : "r"(bar_s), "r"(0)
                         : "memory");
        }
    }
    double2 acc[2][CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[0][j] = acc[1][j] = make_double2(0.0, 0.0);
    if constexpr (PRE) {
        // combine the raw term values ONCE per nonzero (the 8 lanes of a row would each repeat these 2 p DFMA): every thread
        // takes up to 4 nonzeros into registers, then the coefficients c = sum_i c_i a_i are written over the front of the slice
        double2 m[4];
        int cnt = 0;
        for (int e = tid; e < nnz && cnt < 4; e += nth, ++cnt) {
            double v[VW];
#pragma unroll
            for (int u = 0; u < VW; u += 2) {
                const double2 w = *(const double2*)(sA + (size_t)e * VW + u);
                v[u] = w.x;
                v[u + 1] = w.y;
            }
            m[cnt] = combine<VW, false>(v, [&](int i) { return cp.c[i]; });
        }
        __syncthreads();
        cnt = 0;
        for (int e = tid; e < nnz && cnt < 4; e += nth, ++cnt) ((double2*)(smem_raw + oA))[e] = m[cnt];
        __syncthreads();
    }
    auto load_vals_s = [&](int idx, double (&v)[VW]) {
        if constexpr (PRE) {
            const double2 w = ((const double2*)sA)[idx];
            v[0] = w.x;
            v[1] = w.y;
        } else {
            const double* vp = sA + (size_t)idx * VW;
#pragma unroll
            for (int u = 0; u < VW; u += 2) {
                const double2 w = *(const double2*)(vp + u);
                v[u] = w.x;
                v[u + 1] = w.y;
            }
        }
    };
    auto coef_of = [&](const double (&v)[VW]) {
        if constexpr (PRE) return make_double2(v[0], v[1]);
        else return combine<VW, false>(v, [&](int i) { return cp.c[i]; });
    };
    auto load_x = [&](int li, double2 (&x)[CPT], int ncp) {
        const double2* xr = sV + li * kt + gc;
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (j < ncp) x[j] = xr[j * GC];
    };
    auto fma_one = [&](const double (&v)[VW], const double2 (&x)[CPT], double2 (&ac)[CPT], int ncp) {
        const double2 m = coef_of(v);
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (j < ncp) cfma(ac[j], m, x[j]);
    };
    const bool pair = CPT == 3 && kt <= 20;  // two consecutive nonzeros share the half-empty third piece of their V rows
    double2 acc3 = make_double2(0.0, 0.0);
    constexpr int UN = CPT >= 3 ? 2 : 4;
    int base = start;
    if (pair) {
#pragma unroll 2
        for (; base + 2 <= end; base += 2) {
            const int la = (int)sL[base], lb = (int)sL[base + 1];
            double va[VW], vb[VW];
            double2 xa[CPT], xb[CPT];
            load_vals_s(base, va);
            load_vals_s(base + 1, vb);
            load_x(la, xa, 2);
            load_x(lb, xb, 2);
            const double2 x3 = sV[((gc < 4) ? la : lb) * kt + 16 + (gc & 3)];
            const double2 ma = coef_of(va), mb = coef_of(vb);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                cfma(acc[0][j], ma, xa[j]);
                cfma(acc[1][j], mb, xb[j]);
            }
            cfma(acc3, (gc < 4) ? ma : mb, x3);
        }
    } else {
        for (; base + UN <= end; base += UN) {
            int li[UN];
            double v[UN][VW];
            double2 x[UN][CPT];
#pragma unroll
            for (int u = 0; u < UN; ++u) li[u] = (int)sL[base + u];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                load_vals_s(base + u, v[u]);
                load_x(li[u], x[u], CPT);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) fma_one(v[u], x[u], acc[u & 1], CPT);
        }
    }
    for (; base < end; ++base) {
        double v[VW];
        double2 x[CPT];
        load_vals_s(base, v);
        load_x((int)sL[base], x, CPT);
        fma_one(v, x, acc[0], CPT);
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        acc[0][j].x += acc[1][j].x;
        acc[0][j].y += acc[1][j].y;
    }
    if (pair) {
        const double tx = __shfl_down_sync(0xffffffffu, acc3.x, 4), ty = __shfl_down_sync(0xffffffffu, acc3.y, 4);
        if (gc < 4) {
            acc[0][2].x += acc3.x + tx;
            acc[0][2].y += acc3.y + ty;
        }
    }
    if (zrow >= 0) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int c = gc + j * GC;
            if (c < kt) Z[(size_t)zrow * ldz + c] = acc[0][j];
        }
    }
}

// Host side of the two-dimensional tiles: line length from the pattern, tiles = S segments x R rows, per tile the sorted distinct
// columns (runs of consecutive columns = bulk copies), tile-local 16-bit indices stored tile by tile.
static int spmf_build_tiles2d(const nepb_spmf* h) {
    nepb_spmf::TileSet2D& T = h->tiling2d;
    if (T.state) return T.state;
    T.state = -1;
    const int64_t n = h->n;
    const int32_t* rp = h->h_rowptr;
    const int32_t* ci = h->h_colind;
    const int S = std::max(1, std::min(8, getenv("NEPB_SPMM_2D_S") ? atoi(getenv("NEPB_SPMM_2D_S")) : 4));
    const int R = std::max(1, std::min(32, getenv("NEPB_SPMM_2D_R") ? atoi(getenv("NEPB_SPMM_2D_R")) : 8));
    if (S * R * 8 > 256 || n < 64) return -1;
    // dominant column offset ("grid line length") from a sample of rows
    int64_t line = 0;
    {
        std::vector<std::pair<int64_t, int64_t>> offs;
        const int64_t step = std::max<int64_t>(1, n / 4096);
        int64_t total = 0;
        std::vector<int64_t> d;
        for (int64_t r = 0; r < n; r += step)
            for (int32_t e = rp[r]; e < rp[r + 1]; ++e) {
                ++total;
                const int64_t a = std::llabs((long long)ci[e] - (long long)r);
                if (a >= 32) d.push_back(a);
            }
        std::sort(d.begin(), d.end());
        int64_t best = 0;
        for (size_t i = 0; i < d.size();) {
            size_t j = i;
            while (j < d.size() && d[j] == d[i]) ++j;
            if ((int64_t)(j - i) > best) {
                best = (int64_t)(j - i);
                line = d[i];
            }
            i = j;
        }
        if (getenv("NEPB_SPMM_2D_LINE")) line = atoll(getenv("NEPB_SPMM_2D_LINE"));
        else if (best * 25 < total) return -1;  // no offset carries >= 4 % of the nonzeros (a 21-point stencil: 9.5 %): keep the 1D tiles
        if (line < 2 * R || line * S > n) return -1;
    }
    // tiles: whole super-blocks of S lines first, the remaining rows as S consecutive segments
    struct Seg {
        int64_t row0;
        int rows;
    };
    std::vector<Seg> segs;  // S per tile (rows = 0: unused)
    const int64_t sb = (int64_t)S * line, nsb = n / sb;
    for (int64_t b = 0; b < nsb; ++b)
        for (int64_t i0 = 0; i0 < line; i0 += R)
            for (int s = 0; s < S; ++s) segs.push_back({b * sb + s * line + i0, (int)std::min<int64_t>(R, line - i0)});
    for (int64_t r = nsb * sb; r < n; r += (int64_t)S * R)
        for (int s = 0; s < S; ++s) {
            const int64_t r0 = r + (int64_t)s * R;
            segs.push_back({std::min(r0, n), (int)std::max<int64_t>(0, std::min<int64_t>(R, n - r0))});
        }
    const int64_t ntiles = (int64_t)segs.size() / S;
    if (ntiles >= (int64_t)1 << 30) return -1;
    const int DS = 2 + S;
    std::vector<int4> desc((size_t)ntiles * DS);
    constexpr int64_t CHUNK = 512;
    const int64_t nchunks = (ntiles + CHUNK - 1) / CHUNK;
    std::vector<std::vector<int32_t>> ccols(nchunks);
    std::vector<std::vector<int2>> cruns(nchunks);
    std::vector<std::vector<uint16_t>> clidx(nchunks);
    int bad = 0, maxc = 0, maxz = 0;
#pragma omp parallel
    {
        std::vector<int32_t> stamp((size_t)n, -1), local((size_t)n, 0), cur;
#pragma omp for schedule(dynamic, 1) reduction(| : bad) reduction(max : maxc) reduction(max : maxz)
        for (int64_t ch = 0; ch < nchunks; ++ch) {
            for (int64_t t = ch * CHUNK; t < std::min(ntiles, (ch + 1) * CHUNK); ++t) {
                cur.clear();
                int nnz = 0, nseg = 0;
                for (int s = 0; s < S; ++s) {
                    const Seg& g = segs[(size_t)t * S + s];
                    if (g.rows > 0) nseg = s + 1;
                    for (int64_t r = g.row0; r < g.row0 + g.rows; ++r)
                        for (int32_t e = rp[r]; e < rp[r + 1]; ++e) {
                            ++nnz;
                            if (stamp[ci[e]] != (int32_t)t) {
                                stamp[ci[e]] = (int32_t)t;
                                cur.push_back(ci[e]);
                            }
                        }
                }
                if (cur.size() > 4096) bad |= 1;
                std::sort(cur.begin(), cur.end());
                for (size_t u = 0; u < cur.size(); ++u) local[cur[u]] = (int32_t)u;
                const int run_first = (int)cruns[ch].size();
                for (size_t u = 0; u < cur.size();) {
                    size_t v = u + 1;
                    while (v < cur.size() && cur[v] == cur[v - 1] + 1 && v - u < 32767) ++v;
                    cruns[ch].push_back(make_int2(cur[u], (int)u | ((int)(v - u) << 16)));
                    u = v;
                }
                int4* dd = desc.data() + (size_t)t * DS;
                dd[0] = make_int4((int)ccols[ch].size(), (int)cur.size(), run_first, (int)cruns[ch].size() - run_first);
                dd[1] = make_int4((int)clidx[ch].size(), nnz, nseg, 0);
                for (int s = 0; s < S; ++s) {
                    const Seg& g = segs[(size_t)t * S + s];
                    dd[2 + s] = make_int4((int)g.row0, g.rows, g.rows ? rp[g.row0] : 0, g.rows ? rp[g.row0 + g.rows] - rp[g.row0] : 0);
                    for (int64_t r = g.row0; r < g.row0 + g.rows; ++r)
                        for (int32_t e = rp[r]; e < rp[r + 1]; ++e) clidx[ch].push_back((uint16_t)local[ci[e]]);
                }
                while (clidx[ch].size() % 8) clidx[ch].push_back(0);
                ccols[ch].insert(ccols[ch].end(), cur.begin(), cur.end());
                maxc = std::max(maxc, (int)cur.size());
                maxz = std::max(maxz, nnz);
            }
        }
    }
    if (bad) return -1;
    std::vector<int32_t> cols;
    std::vector<int2> runs;
    std::vector<uint16_t> lidx;
    for (int64_t ch = 0; ch < nchunks; ++ch) {
        const int coff = (int)cols.size(), roff = (int)runs.size();
        const size_t loff = lidx.size();
        if (loff + clidx[ch].size() >= (size_t)std::numeric_limits<int32_t>::max() || cols.size() + ccols[ch].size() >= (size_t)std::numeric_limits<int32_t>::max())
            return -1;
        for (int64_t t = ch * CHUNK; t < std::min(ntiles, (ch + 1) * CHUNK); ++t) {
            desc[(size_t)t * DS].x += coff;
            desc[(size_t)t * DS].z += roff;
            desc[(size_t)t * DS + 1].x += (int)loff;
        }
        cols.insert(cols.end(), ccols[ch].begin(), ccols[ch].end());
        runs.insert(runs.end(), cruns[ch].begin(), cruns[ch].end());
        lidx.insert(lidx.end(), clidx[ch].begin(), clidx[ch].end());
    }
    auto up = [&](auto& buf, const auto& vec) {
        cudaError_t e = buf.alloc(std::max<size_t>(vec.size(), 1));
        if (e == cudaSuccess && !vec.empty()) e = cudaMemcpy(buf.p, vec.data(), sizeof(vec[0]) * vec.size(), cudaMemcpyHostToDevice);
        return e;
    };
    cudaError_t e = up(T.desc, desc);
    if (e == cudaSuccess) e = up(T.cols, cols);
    if (e == cudaSuccess) e = up(T.lidx, lidx);
    if (e == cudaSuccess) e = up(T.runs, runs);
    if (e != cudaSuccess) {
        cudaGetLastError();
        T.desc.release();
        T.cols.release();
        T.lidx.release();
        T.runs.release();
        return -1;
    }
    T.S = S;
    T.R = R;
    T.line = (int)line;
    T.ntiles = ntiles;
    T.max_cols = maxc;
    T.max_nnz = maxz;
    T.state = 1;
    return 1;
}

static size_t tma2d_smem_bytes(const nepb_spmf::TileSet2D& T, int kt, int vw) {
    return (size_t)T.max_cols * kt * 16 + (size_t)T.max_nnz * vw * 8 + (size_t)((T.max_nnz + 7) & ~7) * 2 + 16;
}

// returns 1 when the product was launched on the two-dimensional tiles
template <int VW>
static int launch_tma2d_vw(const nepb_spmf* h, const nepb_spmf::TileSet2D& T, int kt, int ldv, int ldz, const double2* V, double2* Z, const CoefP& cp) {
    const size_t smem = tma2d_smem_bytes(T, kt, VW);
    // one combine pass per tile instead of 8 redundant ones per nonzero: measured slower at every width (k = 20: 421 vs 373 us,
    // profiles/r2_spmm_2d.txt), kept selectable (NEPB_SPMM_2D_PRE=1)
    const bool pre = getenv("NEPB_SPMM_2D_PRE") && atoi(getenv("NEPB_SPMM_2D_PRE")) != 0 && T.max_nnz <= 4 * T.S * T.R * 8;
#define NEPB_TMA2D(CPT_)                                                                                                              \
    do {                                                                                                                              \
        static size_t attr_done[16] = {0};                                                                                            \
        int dev = 0;                                                                                                                  \
        cudaGetDevice(&dev);                                                                                                          \
        if (smem > 48 * 1024 && smem > attr_done[dev & 15]) {                                                                         \
            NEPB_CUDA(cudaFuncSetAttribute(spmm_tma2d_kernel<VW, CPT_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            NEPB_CUDA(cudaFuncSetAttribute(spmm_tma2d_kernel<VW, CPT_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_done[dev & 15] = smem;                                                                                               \
        }                                                                                                                             \
        if (pre)                                                                                                                      \
            NEPB_LAUNCH((spmm_tma2d_kernel<VW, CPT_, true>), (unsigned)T.ntiles, T.S * T.R * 8, smem, kt, ldv, ldz, T.max_cols, T.max_nnz, T.S, \
                        T.R, T.desc.p, T.runs.p, T.cols.p, h->d_rowptr.p, T.lidx.p, h->d_vals.p, V, Z, cp);                            \
        else                                                                                                                          \
            NEPB_LAUNCH((spmm_tma2d_kernel<VW, CPT_, false>), (unsigned)T.ntiles, T.S * T.R * 8, smem, kt, ldv, ldz, T.max_cols, T.max_nnz, T.S, \
                        T.R, T.desc.p, T.runs.p, T.cols.p, h->d_rowptr.p, T.lidx.p, h->d_vals.p, V, Z, cp);                            \
    } while (0)
    if (kt <= 8) NEPB_TMA2D(1);
    else if (kt <= 16) NEPB_TMA2D(2);
    else if (kt <= 24) NEPB_TMA2D(3);
    else NEPB_TMA2D(4);
#undef NEPB_TMA2D
    return 1;
}

// the TMA-staged tiled product (default for 5..32 columns when the operands allow bulk copies); 0 = does not apply
static int launch_tma(const nepb_spmf* h, bool diag, int kt, int ldv, int ldz, const double2* V, double2* Z, const CoefP& cp,
                      const double2* cdiag) {
    const bool enabled = !(getenv("NEPB_SPMM_TMA") && atoi(getenv("NEPB_SPMM_TMA")) == 0);  // read per call: tests / tools toggle it
    if (!enabled || kt < 5 || getenv("NEPB_SPMM_CFG") || getenv("NEPB_SPMM_BULK")) return 0;
    {
        const int v = getenv("NEPB_SPMM_DBG") ? atoi(getenv("NEPB_SPMM_DBG")) : 0;  // read per call (tools toggle it)
        static int last = 0;
        if (v != last) {
            cudaMemcpyToSymbol(g_spmm_dbg, &v, sizeof(int));
            last = v;
        }
    }
    if (((uintptr_t)V & 15) != 0) return 0;  // bulk copies need 16-byte aligned sources (rows are ldv*16 bytes apart)
    // two-dimensional tiles when the pattern has a dominant line length (SCALAR mode, real values, even term count)
    const bool want2d = !(getenv("NEPB_SPMM_2D") && atoi(getenv("NEPB_SPMM_2D")) == 0);  // read per call: tests / tools toggle it
    if (want2d && !diag && !h->is_complex && (h->vw == 2 || h->vw == 4) && !getenv("NEPB_SPMM_TILE_ROWS") && !getenv("NEPB_SPMM_GC") &&
        spmf_build_tiles2d(h) == 1 && tma2d_smem_bytes(h->tiling2d, kt, h->vw) <= 200 * 1024) {
        const int rc2 = h->vw == 2 ? launch_tma2d_vw<2>(h, h->tiling2d, kt, ldv, ldz, V, Z, cp) : launch_tma2d_vw<4>(h, h->tiling2d, kt, ldv, ldz, V, Z, cp);
        if (rc2 == 1) {
            NEPB_LAUNCH_CHECK();
            return 1;
        }
        if (rc2 < 0) return rc2;
    }
    // measured on C4 (profiles/r2_spmm_variants.txt): 16-row tiles (twice the resident CTAs = pipeline stages) beat 32-row
    // tiles at every width, and 8 lanes per row beat 4 (fewer wavefronts, but too few warps to hide the latencies)
    int which = 1;
    if (const char* e = getenv("NEPB_SPMM_TILE_ROWS")) which = atoi(e) == 32 ? 0 : 1;
    if (spmf_build_tiles(h, which) != 1) return 0;
    const nepb_spmf::TileSet& T = h->tiling[which];
    const int tile_rows = which == 0 ? 32 : 16;
    int gc = 8;
    if (const char* e = getenv("NEPB_SPMM_GC")) gc = atoi(e) == 4 ? 4 : 8;
    if (tma_smem_layout(T.max_cols, T.max_nnz, kt, h->vw, diag, tile_rows).total > 200 * 1024) return 0;
    const int vw = h->vw;
    const bool ca = h->is_complex;
    int rc = 0;
#define NEPB_VW_TMA(VW_, CA_)                                                                                  \
    if (!rc && vw == VW_ && ca == CA_)                                                                          \
        rc = diag ? launch_tma_vw<VW_, CA_, true>(h, T, tile_rows, gc, kt, ldv, ldz, V, Z, cp, cdiag)            \
                  : launch_tma_vw<VW_, CA_, false>(h, T, tile_rows, gc, kt, ldv, ldz, V, Z, cp, cdiag);
    NEPB_VW_TMA(2, false)
    NEPB_VW_TMA(3, false)
    NEPB_VW_TMA(4, false)
    NEPB_VW_TMA(4, true)
    NEPB_VW_TMA(8, true)
#undef NEPB_VW_TMA
    if (rc == 1) {
        NEPB_LAUNCH_CHECK();
    }
    return rc;
}

// tiled path for 5..32 columns (returns 0 when it does not apply: few columns, unsupported value layout, no tiles)
static int launch_tiled(const nepb_spmf* h, bool diag, int kt, int ldv, int ldz, const double2* V, double2* Z, const CoefP& cp,
                        const double2* cdiag) {
    const bool enabled = !(getenv("NEPB_SPMM_TILED") && atoi(getenv("NEPB_SPMM_TILED")) == 0);
    if (!enabled || kt < 5 || getenv("NEPB_SPMM_CFG")) return 0;
    // Measured on config C4 (profiles/r1_spmm_tiled.txt): at k = 8 the tiled kernel runs at 48 % of the HBM roofline against
    // 33 % untiled; from k ~ 16 on the row products are bound by the shared-memory pipe either way (one 128-byte wavefront per
    // nonzero and 8 columns: 3 per nonzero at k = 20, which alone equals the HBM time) and staging buys nothing, so wide
    // blocks keep the untiled kernels unless NEPB_SPMM_TILE_ROWS forces 32- or 16-row tiles.
    int which = 0;
    if (const char* e = getenv("NEPB_SPMM_TILE_ROWS")) which = atoi(e) == 16 ? 1 : 0;
    else if (kt > 12) return 0;
    if (spmf_build_tiles(h, which) != 1) return 0;
    const nepb_spmf::TileSet& T = h->tiling[which];
    if (tiled_smem_bytes(T, kt, diag ? h->vw : 2) > 200 * 1024) return 0;
    const int threads = which == 0 ? 256 : 128;
    const int vw = h->vw;
    const bool ca = h->is_complex;
    int rc = 0;
#define NEPB_VW_TILED(VW_, CA_)                                                                                           \
    if (!rc && vw == VW_ && ca == CA_)                                                                                    \
        rc = diag ? launch_tiled_vw<VW_, CA_, true>(h, T, threads, kt, ldv, ldz, V, Z, cp, cdiag)                          \
                  : launch_tiled_vw<VW_, CA_, false>(h, T, threads, kt, ldv, ldz, V, Z, cp, cdiag);
    NEPB_VW_TILED(2, false)
    NEPB_VW_TILED(3, false)
    NEPB_VW_TILED(4, false)
    NEPB_VW_TILED(4, true)
    NEPB_VW_TILED(8, true)
#undef NEPB_VW_TILED
    if (rc == 1) {
        NEPB_LAUNCH_CHECK();
    }
    return rc;
}

// one column tile (<= 32 columns) of the SCALAR / DIAG product
static int launch_fused(const nepb_spmf* h, bool diag, int kt, int ldv, int ldz, const double2* V, double2* Z,
                        const CoefP& cp, const double2* cdiag) {
    {
        int rc = launch_tma(h, diag, kt, ldv, ldz, V, Z, cp, cdiag);
        if (rc == 1) return NEPB_OK;
        if (rc != 0) return rc;
        rc = launch_tiled(h, diag, kt, ldv, ldz, V, Z, cp, cdiag);
        if (rc == 1) return NEPB_OK;
        if (rc != 0) return rc;
    }
    TileCfg cfg = default_cfg(kt), ecfg;
    bool tuned = false;
    if (env_cfg(ecfg) && ecfg.gc * ecfg.cpt >= kt) {
        cfg = ecfg;
        tuned = true;
    }
    const int vw = h->vw;
    const bool ca = h->is_complex;
    int done = 0;
#define NEPB_VW_CASE(VW_, CA_)                                                                              \
    if (!done && vw == VW_ && ca == CA_) {                                                                   \
        done = diag ? launch_fused_vw<VW_, CA_, true>(h, cfg, kt, ldv, ldz, V, Z, cp, cdiag, tuned)          \
                    : launch_fused_vw<VW_, CA_, false>(h, cfg, kt, ldv, ldz, V, Z, cp, cdiag, tuned);        \
        if (!done && tuned) {                                                                                \
            cfg = default_cfg(kt);                                                                           \
            done = diag ? launch_fused_vw<VW_, CA_, true>(h, cfg, kt, ldv, ldz, V, Z, cp, cdiag, false)      \
                        : launch_fused_vw<VW_, CA_, false>(h, cfg, kt, ldv, ldz, V, Z, cp, cdiag, false);    \
        }                                                                                                    \
    }
    NEPB_VW_CASE(2, false)
    NEPB_VW_CASE(3, false)
    NEPB_VW_CASE(4, false)
    NEPB_VW_CASE(4, true)
    NEPB_VW_CASE(8, true)
#undef NEPB_VW_CASE
    if (!done) {
        const int n = (int)h->n;
        if (diag)
            NEPB_LAUNCH((spmm_fused_generic_kernel<true>), (n + 31) / 32, 256, 0, n, kt, ldv, ldz, h->d_rowptr.p,
                        h->d_colind.p, h->d_vals.p, V, Z, cp, cdiag, h->p, h->is_complex);
        else
            NEPB_LAUNCH((spmm_fused_generic_kernel<false>), (n + 31) / 32, 256, 0, n, kt, ldv, ldz, h->d_rowptr.p,
                        h->d_colind.p, h->d_vals.p, V, Z, cp, cdiag, h->p, h->is_complex);
    }
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

static int launch_stacked(const nepb_spmf* h, int q, const double2* X, double2* Z, int ldz) {
    const int n = (int)h->n;
    const int vw = h->vw;
    const bool ca = h->is_complex;
    bool done = false;
#define NEPB_ST_CASE(VW_, CA_)                                                                                      \
    if (!done && vw == VW_ && ca == CA_) {                                                                           \
        if (q == 1)                                                                                                  \
            NEPB_LAUNCH((spmm_stacked_kernel<VW_, CA_, 1, 8, 4>), (n + 31) / 32, 256, 0, n, q, h->d_rowptr.p,         \
                        h->d_colind.p, h->d_vals.p, X, Z, ldz);                                                      \
        else                                                                                                         \
            NEPB_LAUNCH((spmm_stacked_kernel<VW_, CA_, 4, 2, 4>), (n + 31) / 32, 256, 0, n, q, h->d_rowptr.p,         \
                        h->d_colind.p, h->d_vals.p, X, Z, ldz);                                                      \
        done = true;                                                                                                 \
    }
    NEPB_ST_CASE(2, false)
    NEPB_ST_CASE(3, false)
    NEPB_ST_CASE(4, false)
    NEPB_ST_CASE(4, true)
    NEPB_ST_CASE(8, true)
#undef NEPB_ST_CASE
    if (!done)
        NEPB_LAUNCH(spmm_stacked_generic_kernel, (n + 31) / 32, 256, 0, n, q, h->p, h->is_complex, h->d_rowptr.p,
                    h->d_colind.p, h->d_vals.p, X, Z, ldz);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

// Z = sum_i A_i (V C_i) with device row-major operands (dV: n x k, ld k ; dZ: n x q, ld q)
int spmf_apply_device_ld(const nepb_spmf* h, int mode, int k, int q, const double2* dV, int ldv, const double* C, double2* dZ, int ldz);
int spmf_apply_device(const nepb_spmf* h, int mode, int k, int q, const double2* dV, const double* C, double2* dZ) {
    return spmf_apply_device_ld(h, mode, k, q, dV, k, C, dZ, q);
}

// operands with explicit leading dimensions (column windows of wider row-major blocks)
int spmf_apply_device_ld(const nepb_spmf* h, int mode, int k, int q, const double2* dV, int ldv, const double* C, double2* dZ, int ldz) {
    NEPB_CHECK_ARG(h && dV && dZ && C, "NULL argument");
    NEPB_CHECK_ARG(k >= 1 && q >= 1, "k and q must be positive (k=%d q=%d)", k, q);
    const int p = h->p;
    if (mode == NEPB_COEF_SCALAR || mode == NEPB_COEF_DIAG) {
        NEPB_CHECK_ARG(q == k, "SCALAR/DIAG coefficient modes need q == k (k=%d q=%d)", k, q);
        NEPB_CHECK_ARG(p <= MAXP, "p=%d exceeds the %d terms supported by the fused kernel", p, MAXP);
        CoefP cp;
        memset(&cp, 0, sizeof(cp));
        const double2* cdiag = nullptr;
        if (mode == NEPB_COEF_SCALAR) {
            for (int i = 0; i < p; ++i) cp.c[i] = make_double2(C[2 * i], C[2 * i + 1]);
        } else {
            NEPB_CUDA(h->d_coef.reserve((size_t)2 * p * k));
            NEPB_CUDA(cudaMemcpyAsync(h->d_coef.p, C, sizeof(double) * 2 * p * k, cudaMemcpyHostToDevice, stream()));
            cdiag = (const double2*)h->d_coef.p;
        }
        for (int k0 = 0; k0 < k; k0 += 32) {
            const int kt = std::min(32, k - k0);
            int rc = launch_fused(h, mode == NEPB_COEF_DIAG, kt, ldv, ldz, dV + k0, dZ + k0, cp, cdiag ? cdiag + (size_t)p * k0 : nullptr);
            if (rc) return rc;
        }
        return NEPB_OK;
    }
    NEPB_CHECK_ARG(mode == NEPB_COEF_GENERAL, "unknown coefficient mode %d", mode);
    // stage 1: Cs[kk][c*p + i] = C_i[kk, c]  (host re-pack, k*q*p complex), X = V * Cs
    const int w = p * q;
    std::vector<double> cs((size_t)2 * k * w);
    for (int i = 0; i < p; ++i)
        for (int c = 0; c < q; ++c)
            for (int kk = 0; kk < k; ++kk) {
                const double* s = C + 2 * ((size_t)i * k * q + (size_t)c * k + kk);
                double* d = cs.data() + 2 * ((size_t)kk * w + (size_t)c * p + i);
                d[0] = s[0];
                d[1] = s[1];
            }
    NEPB_CUDA(h->d_coef.reserve(cs.size()));
    NEPB_CUDA(cudaMemcpyAsync(h->d_coef.p, cs.data(), cs.size() * sizeof(double), cudaMemcpyHostToDevice, stream()));
    NEPB_CUDA(cudaStreamSynchronize(stream()));  // cs is a local buffer
    NEPB_CUDA(h->d_tmp_x.reserve((size_t)2 * h->n * w));
    static const bool v1 = getenv("NEPB_GENERAL_V1") && atoi(getenv("NEPB_GENERAL_V1")) != 0;
    // stage 1: row-streaming panel product when the coefficient block fits shared memory
    if (!v1 && (size_t)k * (w + 1) * 16 <= 96 * 1024) {
        const size_t smem = (size_t)k * (w + 1) * 16;
        static size_t attr_done[16] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (smem > 48 * 1024 && smem > attr_done[dev & 15]) {
            NEPB_CUDA(cudaFuncSetAttribute(panel_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            NEPB_CUDA(cudaFuncSetAttribute(panel_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_done[dev & 15] = 96 * 1024;
        }
        const unsigned grid = (unsigned)((h->n + 255) / 256);
        if (w <= 4)
            NEPB_LAUNCH(panel_rows_kernel<4>, grid, 256, smem, (int)h->n, k, w, dV, ldv, (const double2*)h->d_coef.p, (double2*)h->d_tmp_x.p);
        else
            NEPB_LAUNCH(panel_rows_kernel<8>, grid, 256, smem, (int)h->n, k, w, dV, ldv, (const double2*)h->d_coef.p, (double2*)h->d_tmp_x.p);
    } else {
        NEPB_LAUNCH(panel_gemm_kernel, (int)((h->n + PANEL_R - 1) / PANEL_R), 256, 0, (int)h->n, k, w, dV, ldv,
                    (const double2*)h->d_coef.p, (double2*)h->d_tmp_x.p);
    }
    NEPB_LAUNCH_CHECK();
    // stage 2: TMA-staged stacked gather for real term values with p <= 4 and up to 32 output columns
    if (!v1 && !h->is_complex && p <= 4 && h->vw == p && q <= 32 && spmf_build_tiles(h, 1) == 1) {
        const nepb_spmf::TileSet& T = h->tiling[1];
        const TmaSmem L = tma_smem_layout(T.max_cols, T.max_nnz, w, 4, true, 16);
        if (L.total <= 200 * 1024) {
#define NEPB_STK(QS_, CPT_)                                                                                                          \
    do {                                                                                                                             \
        static size_t attr_done2[16] = {0};                                                                                          \
        int dev2 = 0;                                                                                                                \
        cudaGetDevice(&dev2);                                                                                                        \
        if (L.total > 48 * 1024 && L.total > attr_done2[dev2 & 15]) {                                                                \
            NEPB_CUDA(cudaFuncSetAttribute(spmm_stacked_tma_kernel<QS_, CPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total)); \
            attr_done2[dev2 & 15] = L.total;                                                                                         \
        }                                                                                                                            \
        NEPB_LAUNCH((spmm_stacked_tma_kernel<QS_, CPT_>), (unsigned)T.ntiles, 16 * 4 * QS_, L.total, q, p, ldz, T.max_cols, T.max_nnz, 16, \
                    T.tiles.p, T.runs.p, h->d_rowptr.p, T.lidx.p, h->d_vals.p, (const double2*)h->d_tmp_x.p, dZ);                     \
    } while (0)
            if (q == 1) NEPB_STK(1, 1);
            else if (q <= 8) NEPB_STK(2, 4);
            else if (q <= 16) NEPB_STK(2, 8);
            else NEPB_STK(2, 16);
#undef NEPB_STK
            NEPB_LAUNCH_CHECK();
            return NEPB_OK;
        }
    }
    return launch_stacked(h, q, (const double2*)h->d_tmp_x.p, dZ, ldz);
}

int upload_colmajor(int64_t n, int kc, const double* host, int64_t ld, DevBuf<double>& stage, double* dst, int ldd, int k0);    // hostcopy.cu
int download_colmajor(int64_t n, int kc, const double* src, int lds, int k0, DevBuf<double>& stage, double* host, int64_t ld);

}  // namespace nepb

nepb_spmf::~nepb_spmf() {
    free(h_colptr);
    free(h_rowval);
    free(h_rowptr);
    free(h_colind);
    free(h_csr_of_csc);
}

using namespace nepb;

namespace nepb {
void lu_symbolic_release(void* p);
}

extern "C" {

int nepb_spmf_create(int64_t n, int p, const int64_t* const* colptr, const int64_t* const* rowval,
                     const void* const* nzval, int val_is_complex, int index_base, nepb_spmf** out) {
    NEPB_CHECK_ARG(out, "out is NULL");
    *out = nullptr;
    NEPB_CHECK_ARG(n >= 1, "n must be positive (n=%lld)", (long long)n);
    NEPB_CHECK_ARG(p >= 1, "an SPMF needs at least one term (p=%d)", p);
    NEPB_CHECK_ARG(index_base == 0 || index_base == 1, "index_base must be 0 or 1");
    NEPB_CHECK_ARG(colptr && rowval && nzval, "NULL array of arrays");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libnepb200 has no CPU fallback");
        return NEPB_E_CUDA;
    }
    UnionCSR u;
    int rc = build_union_csr(n, p, colptr, rowval, nzval, val_is_complex, index_base, u);
    if (rc) return rc;
    nepb_spmf* h = new nepb_spmf();
    h->n = n;
    h->p = p;
    h->nnz = u.nnz;
    h->is_complex = val_is_complex ? 1 : 0;
    h->index_base = index_base;
    h->vw = u.vw;
    auto dup = [](const void* src, size_t bytes) {
        void* d = malloc(bytes ? bytes : 1);
        if (d && bytes) memcpy(d, src, bytes);
        return d;
    };
    h->h_colptr = (int64_t*)dup(u.colptr.data(), sizeof(int64_t) * (n + 1));
    h->h_rowval = (int32_t*)dup(u.rowval.data(), sizeof(int32_t) * u.nnz);
    h->h_rowptr = (int32_t*)dup(u.rowptr.data(), sizeof(int32_t) * (n + 1));
    h->h_colind = (int32_t*)dup(u.colind.data(), sizeof(int32_t) * u.nnz);
    h->h_csr_of_csc = (int32_t*)dup(u.csr_of_csc.data(), sizeof(int32_t) * u.nnz);
#define NEPB_UP(buf, src, count, T)                                                                       \
    do {                                                                                                  \
        cudaError_t e_ = buf.alloc((count) + 16 / sizeof(T)); /* slack: bulk copies round slice ends up to 16 bytes */ \
        if (e_ == cudaSuccess) e_ = cudaMemcpy(buf.p, src, sizeof(T) * (count), cudaMemcpyHostToDevice);  \
        if (e_ != cudaSuccess) {                                                                          \
            set_error("CUDA error while uploading the operator: %s", cudaGetErrorString(e_));             \
            delete h;                                                                                     \
            return e_ == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;                          \
        }                                                                                                 \
    } while (0)
    NEPB_UP(h->d_rowptr, u.rowptr.data(), (size_t)n + 1, int32_t);
    NEPB_UP(h->d_colind, u.colind.data(), (size_t)u.nnz, int32_t);
    NEPB_UP(h->d_csr_of_csc, u.csr_of_csc.data(), (size_t)u.nnz, int32_t);
    NEPB_UP(h->d_vals, u.vals.data(), (size_t)u.nnz * u.vw, double);
#undef NEPB_UP
    *out = h;
    return NEPB_OK;
}

int nepb_spmf_destroy(nepb_spmf* h) {
    if (h) {
        if (h->lu_symbolic) lu_symbolic_release(h->lu_symbolic);
        for (void* m : h->lu_matched) lu_symbolic_release(m);
        delete h;
    }
    return NEPB_OK;
}

int nepb_spmf_info(const nepb_spmf* h, int64_t* n, int* p, int64_t* nnz_union, int* val_is_complex) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    if (n) *n = h->n;
    if (p) *p = h->p;
    if (nnz_union) *nnz_union = h->nnz;
    if (val_is_complex) *val_is_complex = h->is_complex;
    return NEPB_OK;
}

int nepb_spmf_tiles_info(const nepb_spmf* h, int64_t* ntiles, int64_t* distinct_total, int* max_distinct) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    const int st = spmf_build_tiles(h, 0);
    if (ntiles) *ntiles = st == 1 ? h->tiling[0].ntiles : 0;
    if (distinct_total) *distinct_total = st == 1 ? h->tiling[0].cols_total : 0;
    if (max_distinct) *max_distinct = st == 1 ? h->tiling[0].max_cols : 0;
    return NEPB_OK;
}

// two-dimensional tiles of the multi-column product: line length found in the pattern (0: none, the 1D tiles are used), segments
// per tile, rows per segment, tiles, staged rows of V in total (distinct columns summed over the tiles)
int nepb_spmf_tiles2d_info(const nepb_spmf* h, int* line, int* segments, int* seg_rows, int64_t* ntiles, int64_t* distinct_total) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    const bool ok = spmf_build_tiles2d(h) == 1;
    const nepb_spmf::TileSet2D& T = h->tiling2d;
    if (line) *line = ok ? T.line : 0;
    if (segments) *segments = ok ? T.S : 0;
    if (seg_rows) *seg_rows = ok ? T.R : 0;
    if (ntiles) *ntiles = ok ? T.ntiles : 0;
    if (distinct_total) *distinct_total = ok ? (int64_t)(T.cols.n) : 0;
    return NEPB_OK;
}

int nepb_spmf_pattern(const nepb_spmf* h, int64_t* colptr, int64_t* rowval) {
    NEPB_CHECK_ARG(h && colptr && rowval, "NULL argument");
    for (int64_t j = 0; j <= h->n; ++j) colptr[j] = h->h_colptr[j] + h->index_base;
    for (int64_t e = 0; e < h->nnz; ++e) rowval[e] = (int64_t)h->h_rowval[e] + h->index_base;
    return NEPB_OK;
}

int nepb_spmf_pattern_csr(const nepb_spmf* h, int32_t* rowptr, int32_t* colind, int32_t* csr_of_csc) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    if (rowptr) memcpy(rowptr, h->h_rowptr, sizeof(int32_t) * (h->n + 1));
    if (colind) memcpy(colind, h->h_colind, sizeof(int32_t) * h->nnz);
    if (csr_of_csc) memcpy(csr_of_csc, h->h_csr_of_csc, sizeof(int32_t) * h->nnz);
    return NEPB_OK;
}

int nepb_spmf_mder(const nepb_spmf* h, const double* coef, double* nzval_out) {
    NEPB_CHECK_ARG(h && coef && nzval_out, "NULL argument");
    NEPB_CHECK_ARG(h->p <= MAXP, "p=%d exceeds %d", h->p, MAXP);
    CoefP cp;
    memset(&cp, 0, sizeof(cp));
    for (int i = 0; i < h->p; ++i) cp.c[i] = make_double2(coef[2 * i], coef[2 * i + 1]);
    NEPB_CUDA(h->d_tmp_out.reserve((size_t)2 * h->nnz));
    NEPB_LAUNCH(mder_kernel, (unsigned)((h->nnz + 255) / 256), 256, 0, h->nnz, h->p, h->is_complex, h->d_csr_of_csc.p,
                h->d_vals.p, cp, (double2*)h->d_tmp_out.p);
    NEPB_LAUNCH_CHECK();
    NEPB_CUDA(cudaMemcpyAsync(nzval_out, h->d_tmp_out.p, sizeof(double) * 2 * h->nnz, cudaMemcpyDeviceToHost, stream()));
    NEPB_CUDA(cudaStreamSynchronize(stream()));
    return NEPB_OK;
}

int nepb_spmf_apply(const nepb_spmf* h, int mode, int k, int q, const double* V, int64_t ldv, const double* C,
                    double* Z, int64_t ldz) {
    NEPB_CHECK_ARG(h && V && C && Z, "NULL argument");
    NEPB_CHECK_ARG(k >= 1 && q >= 1, "k and q must be positive");
    NEPB_CHECK_ARG(ldv >= h->n && ldz >= h->n, "leading dimension smaller than n");
    const int64_t n = h->n;
    NEPB_CUDA(h->d_tmp_in.reserve((size_t)2 * n * k));
    NEPB_CUDA(h->d_tmp_out.reserve((size_t)2 * n * q));
    int rc = upload_colmajor(n, k, V, ldv, h->d_stage, h->d_tmp_in.p, k, 0);
    if (rc) return rc;
    rc = spmf_apply_device(h, mode, k, q, (const double2*)h->d_tmp_in.p, C, (double2*)h->d_tmp_out.p);
    if (rc) return rc;
    return download_colmajor(n, q, h->d_tmp_out.p, q, 0, h->d_stage, Z, ldz);
}

int nepb_block_create(int64_t n, int k, nepb_block** out) {
    NEPB_CHECK_ARG(out && n >= 1 && k >= 1, "bad arguments");
    nepb_block* b = new nepb_block();
    b->n = n;
    b->k = k;
    cudaError_t e = b->d.alloc((size_t)2 * n * k);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d.p, 0, sizeof(double) * 2 * n * k, stream());
    if (e != cudaSuccess) {
        set_error("cudaMalloc of a %lld x %d block failed: %s", (long long)n, k, cudaGetErrorString(e));
        delete b;
        return e == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;
    }
    *out = b;
    return NEPB_OK;
}

int nepb_block_destroy(nepb_block* b) {
    delete b;
    return NEPB_OK;
}

static DevBuf<double> g_block_stage;

int nepb_block_upload(nepb_block* b, int k0, int kc, const double* host, int64_t ld) {
    NEPB_CHECK_ARG(b && host && k0 >= 0 && kc >= 1 && k0 + kc <= b->k && ld >= b->n, "bad arguments");
    return upload_colmajor(b->n, kc, host, ld, g_block_stage, b->d.p, b->k, k0);
}

int nepb_block_download(const nepb_block* b, int k0, int kc, double* host, int64_t ld) {
    NEPB_CHECK_ARG(b && host && k0 >= 0 && kc >= 1 && k0 + kc <= b->k && ld >= b->n, "bad arguments");
    return download_colmajor(b->n, kc, b->d.p, b->k, k0, g_block_stage, host, ld);
}

void* nepb_block_dev_ptr(nepb_block* b) { return b ? (void*)b->d.p : nullptr; }

int nepb_spmf_apply_block(const nepb_spmf* h, int mode, const nepb_block* V, int q, const double* C, nepb_block* Z) {
    NEPB_CHECK_ARG(h && V && Z && C, "NULL argument");
    NEPB_CHECK_ARG(V->n == h->n && Z->n == h->n, "block row count differs from the operator size");
    NEPB_CHECK_ARG(Z->k == q, "Z has %d columns, expected q=%d", Z->k, q);
    NEPB_CHECK_ARG(V->d.p != Z->d.p, "V and Z must not alias");
    return spmf_apply_device(h, mode, V->k, q, (const double2*)V->d.p, C, (double2*)Z->d.p);
}

int nepb_spmf_apply_block_ex(const nepb_spmf* h, int mode, const nepb_block* V, int vcol0, int k, int q, const double* C, nepb_block* Z,
                             int zcol0) {
    NEPB_CHECK_ARG(h && V && Z && C, "NULL argument");
    NEPB_CHECK_ARG(V->n == h->n && Z->n == h->n, "block row count differs from the operator size");
    NEPB_CHECK_ARG(vcol0 >= 0 && k >= 1 && vcol0 + k <= V->k && zcol0 >= 0 && q >= 1 && zcol0 + q <= Z->k, "bad column windows");
    NEPB_CHECK_ARG(!(V == Z && zcol0 < vcol0 + k && vcol0 < zcol0 + q), "input and output columns overlap");
    return spmf_apply_device_ld(h, mode, k, q, (const double2*)V->d.p + vcol0, V->k, C, (double2*)Z->d.p + zcol0, Z->k);
}

int64_t nepb_spmf_apply_bytes(const nepb_spmf* h, int mode, int k, int q) {
    if (!h) return 0;
    const int64_t sA = h->is_complex ? 16 : 8;
    int64_t b = h->nnz * ((int64_t)h->p * sA + 4) + 4 * (h->n + 1) + h->n * (int64_t)k * 16 + h->n * (int64_t)q * 16;
    // GENERAL is two-stage: X = V*[C_1..C_p] (n x p*q) is written once and gathered once
    if (mode == NEPB_COEF_GENERAL) b += 2 * h->n * (int64_t)q * h->p * 16;
    return b;
}

}  // extern "C"
