// WEP-native path: the waveguide eigenvalue problem in its own format, WEP_FD
// (reference: src/gallery_extra/waveguide/Waveguide.jl:203-240), without ever forming a matrix:
//   compute_Mlincomb(::WEP_FD, lambda, V, a)   Waveguide.jl:324-379
//   SchurMatVec * v                            Waveguide.jl:393-402
//   Pinv(nep, lambda, x)                       Waveguide.jl:160-163 (with R / Rinv :165-171)
//
// Interior (nx*nz rows): the Sylvester form A(lambda) X + X B + K .* X of the reference IS a five-point stencil on the
// nz x nx grid (periodic in z) with the variable coefficient K, so it is one streaming pass: 16 B read of X, 16 B of K and
// 16 B written per grid point (the neighbours come out of L1 / L2), HBM-bound.  The derivative terms A'(lambda) X_1 and
// A''(lambda) X_2 ride in the same pass.
// Boundary (2 nz rows): y2 = R( sum_j c_j .* Rinv(v2_j) ), R / Rinv = scaled, reversed DFTs of ODD length nz (3*5*7*k in the
// reference's configurations).  The whole boundary block touches 2 nz (na + 1) numbers, so it is done as a direct
// O(nz^2) transform with an exact twiddle table (index m k mod nz in integer arithmetic): for nz = 945 that is 0.9 M
// complex multiply-adds per transform -- microseconds on the FP64 pipe, no FFT plan, no library, and an error of a few ulp
// (16 interleaved partial sums per output, then a tree).  Sums have a fixed order: results are bitwise reproducible.
#include "common.h"

#include <cmath>
#include <stdlib.h>

using namespace nepb;

struct nepb_wep {
    int nx = 0, nz = 0;
    double hx = 0, hz = 0;
    double kbar_re = 0, kbar_im = 0;
    DevBuf<double> K;    // nz*nx complex, K - k_bar, index z + nz*x (the reference's vec(K))
    DevBuf<double> tw;   // nz complex: exp(2 pi i j / nz)
    DevBuf<double> bb;   // nz complex: the reference's bb (|bb| = 1, bbinv = conj(bb))
    // chirp-z (Bluestein) transforms: FFT length fl = 2^flog >= 2 nz - 1, twiddles exp(-2 pi i j / fl), chirp exp(-i pi k^2 / nz),
    // the filter spectra of both signs in bit-reversed order
    int fl = 0, flog = 0;
    DevBuf<double> ftw, chirp, filt_f, filt_i;
    mutable DevBuf<double> G;
    DevBuf<double> table;  // 2 nz x table_cols complex, row-major: boundary derivative table of one lambda (nepb_wep_set_table)
    int table_cols = 0;
    mutable DevBuf<double> coef, t, u, w, avec;
};

namespace {

struct c2 {
    double re, im;
};
__device__ __forceinline__ c2 ld(const double* p) {
    double2 v = *reinterpret_cast<const double2*>(p);
    return {v.x, v.y};
}
__device__ __forceinline__ void st(double* p, c2 v) { *reinterpret_cast<double2*>(p) = make_double2(v.re, v.im); }
__device__ __forceinline__ c2 mul(c2 a, c2 b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ c2 add(c2 a, c2 b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ c2 sub(c2 a, c2 b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ c2 scal(double s, c2 a) { return {s * a.re, s * a.im}; }
__device__ __forceinline__ c2 fma2(c2 a, c2 b, c2 acc) {
    return {fma(a.re, b.re, fma(-a.im, b.im, acc.re)), fma(a.re, b.im, fma(a.im, b.re, acc.im))};
}

struct InteriorArgs {
    int nx, nz, na;            // na: number of derivative columns used (1..3)
    double ihx2, ihz2, ihz;    // 1/hx^2, 1/hz^2, 1/(2 hz)
    c2 lam, lam2k;             // lambda, lambda^2 + k_bar
    c2 a0, a1, a2;             // coefficients of the columns
    c2 c1s;                    // factor of the C1 term (a0/hx^2 for Mlincomb, -1/hx^2 for the Schur product)
};

// One grid point per thread; z is the fast index, so a warp reads 32 consecutive rows of V.
__global__ void __launch_bounds__(256) wep_interior_kernel(InteriorArgs p, const double* __restrict__ K, const double* __restrict__ V,
                                                           int64_t ldv, const double* __restrict__ v2, int64_t ldv2,
                                                           double* __restrict__ Z, int64_t ldz) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t m = (int64_t)p.nx * p.nz;
    if (i >= m) return;
    int z = (int)(i % p.nz), x = (int)(i / p.nz);
    int64_t iu = (z + 1 < p.nz) ? i + 1 : i + 1 - p.nz;  // periodic in z
    int64_t id = (z > 0) ? i - 1 : i - 1 + p.nz;
    c2 xc = ld(V + 2 * i * ldv), xu = ld(V + 2 * iu * ldv), xd = ld(V + 2 * id * ldv);
    c2 xl = (x > 0) ? ld(V + 2 * (i - p.nz) * ldv) : c2{0, 0};
    c2 xr = (x + 1 < p.nx) ? ld(V + 2 * (i + p.nz) * ldv) : c2{0, 0};
    // A(lambda) X = Dzz X + 2 lambda Dz X + (lambda^2 + k_bar) X ; X B = X Dxx ; K .* X
    c2 acc = scal(p.ihz2, sub(add(xu, xd), scal(2.0, xc)));
    acc = fma2(scal(2.0, p.lam), scal(p.ihz, sub(xu, xd)), acc);
    acc = fma2(p.lam2k, xc, acc);
    acc = add(acc, scal(p.ihx2, sub(add(xl, xr), scal(2.0, xc))));
    acc = fma2(ld(K + 2 * i), xc, acc);
    acc = mul(p.a0, acc);
    if (p.na > 1) {  // A'(lambda) X_1 = 2 Dz X_1 + 2 lambda X_1
        c2 yc = ld(V + 2 * i * ldv + 2), yu = ld(V + 2 * iu * ldv + 2), yd = ld(V + 2 * id * ldv + 2);
        c2 d = add(scal(2.0 * p.ihz, sub(yu, yd)), mul(scal(2.0, p.lam), yc));
        acc = fma2(p.a1, d, acc);
    }
    if (p.na > 2) acc = fma2(p.a2, scal(2.0, ld(V + 2 * i * ldv + 4)), acc);  // A''(lambda) X_2 = 2 X_2
    if (v2) {  // C1 v2: first and last grid column (waveguide_FD.jl:41-49)
        if (x == 0) acc = fma2(p.c1s, ld(v2 + 2 * (int64_t)z * ldv2), acc);
        if (x == p.nx - 1) acc = fma2(p.c1s, ld(v2 + 2 * (int64_t)(p.nz + z) * ldv2), acc);
    }
    st(Z + 2 * i * ldz, acc);
}

// The boundary transforms as direct odd-length DFTs, organised like a small GEMM: a CTA owns WB_M outputs, its 16 x 16 threads
// are (output, k-slice); a thread walks its k-slice with the twiddle index (m k mod nz) kept incrementally in integers,
// multiplies every input row by the twiddle once and feeds up to WB_J columns from registers.  The 16 k-slices of an output are
// folded with a fixed-order shuffle tree (bitwise reproducible).  The inputs (2 nz (na + 1) numbers) stay in L1 / L2.
constexpr int WB_M = 16, WB_S = 16, WB_J = 8;

__device__ __forceinline__ c2 slice_sum(c2 v) {  // sum over the 16 lanes of a half-warp that share an output
    for (int o = 8; o; o >>= 1) {
        v.re += __shfl_xor_sync(0xffffffffu, v.re, o);
        v.im += __shfl_xor_sync(0xffffffffu, v.im, o);
    }
    return v;
}

// t[m] = (1/nz) sum_k W^{+mm k} conj(bb[k]) * ( sum_jj coef[m, jj] v2[h nz + nz-1-k, jj] ),  m = h nz + mm
// = the reference's  sum_jj coef[:, jj] .* [Rinv(v2_jj[1:nz]); Rinv(v2_jj[nz+1:2nz])]   (Waveguide.jl:366-373, Rinv :169-171)
__global__ void __launch_bounds__(WB_M* WB_S) wep_boundary_inv_kernel(int nz, int na, const double* __restrict__ tw,
                                                                      const double* __restrict__ bb, const double* __restrict__ coef, int ldc,
                                                                      const double* __restrict__ avec, const double* __restrict__ v2,
                                                                      int64_t ldv2, double* __restrict__ t) {
    const int sl = threadIdx.x % WB_S, mi = threadIdx.x / WB_S;
    const int tiles = (nz + WB_M - 1) / WB_M;
    const int h = blockIdx.x / tiles, mm = (blockIdx.x % tiles) * WB_M + mi;
    const bool live = mm < nz;
    const int mmc = live ? mm : 0;
    const double* cf = coef + 2 * (int64_t)(h * nz + mmc) * ldc;  // avec != NULL: the table holds D, the coefficient is a_j D[m, j]
    const int step = (int)(((int64_t)mmc * WB_S) % nz);
    c2 tsum{0, 0};
    for (int j0 = 0; j0 < na; j0 += WB_J) {
        const int jc = min(WB_J, na - j0);
        c2 acc[WB_J];
#pragma unroll
        for (int j = 0; j < WB_J; ++j) acc[j] = c2{0, 0};
        int idx = (int)(((int64_t)mmc * sl) % nz);
        for (int k = sl; k < nz; k += WB_S) {
            const c2 b = ld(bb + 2 * k);
            const c2 f = mul(ld(tw + 2 * idx), c2{b.re, -b.im});
            const double* row = v2 + 2 * (int64_t)(h * nz + nz - 1 - k) * ldv2 + 2 * j0;
#pragma unroll
            for (int j = 0; j < WB_J; ++j)
                if (j < jc) acc[j] = fma2(f, ld(row + 2 * j), acc[j]);
            idx += step;
            if (idx >= nz) idx -= nz;
        }
#pragma unroll
        for (int j = 0; j < WB_J; ++j)
            if (j < jc) tsum = fma2(avec ? mul(ld(cf + 2 * (j0 + j)), ld(avec + 2 * (j0 + j))) : ld(cf + 2 * (j0 + j)), acc[j], tsum);
    }
    tsum = slice_sum(tsum);
    if (live && sl == 0) st(t + 2 * (h * nz + mm), scal(1.0 / nz, tsum));
}

// y[h nz + nz-1-j] = bb[j] sum_m t[h nz + m] W^{-j m}  (+ the C2T rows: d1 * first / last grid column + d2 * its neighbour)
// = R(t[1:nz]), R(t[nz+1:2nz])   (Waveguide.jl:374-376, R :165-167; C2T waveguide_FD.jl:52-60)
__global__ void __launch_bounds__(WB_M* WB_S) wep_boundary_fwd_kernel(int nx, int nz, const double* __restrict__ tw,
                                                                      const double* __restrict__ bb, const double* __restrict__ t,
                                                                      const double* __restrict__ V1, int64_t ldv, c2 cd1, c2 cd2,
                                                                      double* __restrict__ y, int64_t ldy) {
    const int sl = threadIdx.x % WB_S, mi = threadIdx.x / WB_S;
    const int tiles = (nz + WB_M - 1) / WB_M;
    const int h = blockIdx.x / tiles, j = (blockIdx.x % tiles) * WB_M + mi;
    const bool live = j < nz;
    const int jc = live ? j : 0;
    const int step = (int)(((int64_t)jc * WB_S) % nz);
    int idx = (int)(((int64_t)jc * sl) % nz);
    c2 acc{0, 0};
    for (int m = sl; m < nz; m += WB_S) {
        const c2 w = ld(tw + 2 * idx);
        acc = fma2(c2{w.re, -w.im}, ld(t + 2 * (h * nz + m)), acc);
        idx += step;
        if (idx >= nz) idx -= nz;
    }
    acc = slice_sum(acc);
    if (live && sl == 0) {
        const int zr = nz - 1 - j;
        c2 out = mul(ld(bb + 2 * j), acc);
        if (V1) {
            int64_t c0 = (h == 0) ? zr : zr + (int64_t)nz * (nx - 1);
            int64_t c1 = (h == 0) ? zr + nz : zr + (int64_t)nz * (nx - 2);
            out = fma2(cd1, ld(V1 + 2 * c0 * ldv), out);
            out = fma2(cd2, ld(V1 + 2 * c1 * ldv), out);
        }
        st(y + 2 * (int64_t)(h * nz + zr) * ldy, out);
    }
}

// ---- chirp-z transforms -------------------------------------------------------------------------------------------
// An odd-length DFT as a convolution (Bluestein): m k = (m^2 + k^2 - (m - k)^2) / 2, so with c[k] = exp(-i pi k^2 / nz)
//   X[m] = sum_k x[k] W^{m k} = c[m] * sum_k (x[k] c[k]) conj(c)[m - k],    W = exp(-2 pi i / nz)
// (conjugate chirps for the inverse sign).  The convolution runs through a power-of-two FFT of length fl >= 2 nz - 1 in shared
// memory: decimation in frequency forwards (natural order in, bit-reversed out), pointwise product with the precomputed filter
// spectrum stored bit-reversed, decimation in time backwards (bit-reversed in, natural out) -- no reordering pass.  Chirp,
// twiddles and filter spectra come from the host in extended precision; k^2 mod 2 nz is exact integer arithmetic.
// One CTA per transform: O(fl log fl) instead of O(nz^2), the 2 na inverse transforms of a compute_Mlincomb call run side by side.
__device__ __forceinline__ void fft_convolve(c2* sm, int fl, const double* __restrict__ ftw_g, const double* __restrict__ filt) {
    const int tid = threadIdx.x, nt = blockDim.x, hl = fl >> 1;
    // twiddles next to the data, one compact run per stage (entry half + pos = exp(-i pi pos / half)): consecutive lanes read
    // consecutive entries in every stage (a single strided table gives 32-way bank conflicts in the middle stages)
    double* ftw = reinterpret_cast<double*>(sm + fl);
    for (int i = tid; i < fl; i += nt) st(ftw + 2 * i, ld(ftw_g + 2 * i));
    for (int half = hl; half >= 1; half >>= 1) {  // DIF
        __syncthreads();
        for (int i = tid; i < hl; i += nt) {
            const int pos = i & (half - 1), i0 = ((i - pos) << 1) + pos, i1 = i0 + half;
            const c2 u = sm[i0], v = sm[i1];
            sm[i0] = add(u, v);
            sm[i1] = mul(sub(u, v), ld(ftw + 2 * (half + pos)));
        }
    }
    __syncthreads();
    for (int i = tid; i < fl; i += nt) sm[i] = mul(sm[i], ld(filt + 2 * i));
    for (int half = 1; half < fl; half <<= 1) {  // DIT with conjugate twiddles (inverse, unscaled)
        __syncthreads();
        for (int i = tid; i < hl; i += nt) {
            const int pos = i & (half - 1), i0 = ((i - pos) << 1) + pos, i1 = i0 + half;
            const c2 w = ld(ftw + 2 * (half + pos));
            const c2 u = sm[i0], v = mul(sm[i1], c2{w.re, -w.im});
            sm[i0] = add(u, v);
            sm[i1] = sub(u, v);
        }
    }
    __syncthreads();
}

// G[(h nz + m) * na + jj] = Rinv(v2[h nz .. , jj])[m]:  block (jj, h)
__global__ void __launch_bounds__(512) wep_czt_inv_kernel(int nz, int na, int fl, const double* __restrict__ ftw, const double* __restrict__ chirp,
                                                          const double* __restrict__ filt, const double* __restrict__ bb,
                                                          const double* __restrict__ v2, int64_t ldv2, double* __restrict__ G) {
    extern __shared__ double2 czt_sm[];
    c2* sm = reinterpret_cast<c2*>(czt_sm);
    const int jj = blockIdx.x, h = blockIdx.y;
    for (int k = threadIdx.x; k < fl; k += blockDim.x) {
        c2 v{0, 0};
        if (k < nz) {
            const c2 b = ld(bb + 2 * k), c = ld(chirp + 2 * k);
            v = mul(mul(ld(v2 + 2 * (int64_t)(h * nz + nz - 1 - k) * ldv2 + 2 * jj), c2{b.re, -b.im}), c2{c.re, -c.im});
        }
        sm[k] = v;
    }
    fft_convolve(sm, fl, ftw, filt);
    const double sc = 1.0 / ((double)fl * (double)nz);
    for (int m = threadIdx.x; m < nz; m += blockDim.x) {
        const c2 c = ld(chirp + 2 * m);
        st(G + 2 * ((int64_t)(h * nz + m) * na + jj), scal(sc, mul(sm[m], c2{c.re, -c.im})));
    }
}

// t[m] = chirp[mm] * sum_jj coef[m, jj] (a_jj) G[m, jj], m = h nz + mm: one warp per output, lanes along the columns (both rows are
// contiguous), fixed-order shuffle tree.  Used for more than 4 columns: inside the two CTAs of the forward transform the same sums
// are a chain of dependent L2 round trips (measured 30 us at 20 columns, 120 us at 60).
__global__ void __launch_bounds__(256) wep_combine_kernel(int nz, int na, const double* __restrict__ chirp, const double* __restrict__ coef, int ldc,
                                                          const double* __restrict__ avec, const double* __restrict__ G, double* __restrict__ t) {
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= 2 * nz) return;
    const double* cf = coef + 2 * (int64_t)m * ldc;
    const double* g = G + 2 * (int64_t)m * na;
    c2 acc{0, 0};
    for (int jj = lane; jj < na; jj += 32) acc = fma2(avec ? mul(ld(cf + 2 * jj), ld(avec + 2 * jj)) : ld(cf + 2 * jj), ld(g + 2 * jj), acc);
    for (int o = 16; o; o >>= 1) {
        acc.re += __shfl_xor_sync(0xffffffffu, acc.re, o);
        acc.im += __shfl_xor_sync(0xffffffffu, acc.im, o);
    }
    if (lane == 0) st(t + 2 * m, mul(acc, ld(chirp + 2 * (m % nz))));
}

// y[h nz + nz-1-j] = bb[j] * DFT( t )[j] (+ C2T rows), t[m] = sum_jj coef[m, jj] (a_jj) G[m, jj]:  block h
__global__ void __launch_bounds__(512) wep_czt_fwd_kernel(int nx, int nz, int na, int fl, const double* __restrict__ ftw,
                                                          const double* __restrict__ chirp, const double* __restrict__ filt,
                                                          const double* __restrict__ bb, const double* __restrict__ coef, int ldc,
                                                          const double* __restrict__ avec, const double* __restrict__ G,
                                                          const double* __restrict__ tvec, const double* __restrict__ V1, int64_t ldv, c2 cd1,
                                                          c2 cd2, double* __restrict__ y, int64_t ldy) {
    extern __shared__ double2 czt_sm[];
    c2* sm = reinterpret_cast<c2*>(czt_sm);
    const int h = blockIdx.x;
    if (tvec) {  // combined (and chirped) by wep_combine_kernel
        for (int m = threadIdx.x; m < fl; m += blockDim.x) sm[m] = (m < nz) ? ld(tvec + 2 * (h * nz + m)) : c2{0, 0};
    } else if (na <= 4) {  // few columns: one thread per m
        for (int m = threadIdx.x; m < fl; m += blockDim.x) {
            c2 t{0, 0};
            if (m < nz) {
                const double* cf = coef + 2 * (int64_t)(h * nz + m) * ldc;
                const double* g = G + 2 * (int64_t)(h * nz + m) * na;
                for (int jj = 0; jj < na; ++jj) t = fma2(avec ? mul(ld(cf + 2 * jj), ld(avec + 2 * jj)) : ld(cf + 2 * jj), ld(g + 2 * jj), t);
                t = mul(t, ld(chirp + 2 * m));
            }
            sm[m] = t;
        }
    } else {
        // t[m] = sum_jj coef[m, jj] G[m, jj]: 8 lanes per m (lane l takes jj = l, l + 8, ...), folded in a fixed order; the
        // padding m >= nz is zeroed separately so that the loop (whose loads are dependent L2 round trips) only covers nz outputs
        for (int m = nz + threadIdx.x; m < fl; m += blockDim.x) sm[m] = c2{0, 0};
        const int per = blockDim.x / 8, l = threadIdx.x % 8, mo = threadIdx.x / 8;
#pragma unroll 4
        for (int m0 = 0; m0 < nz; m0 += per) {
            const int m = m0 + mo;
            c2 t{0, 0};
            if (m < nz) {
                const double* cf = coef + 2 * (int64_t)(h * nz + m) * ldc;
                const double* g = G + 2 * (int64_t)(h * nz + m) * na;
                for (int jj = l; jj < na; jj += 8) t = fma2(avec ? mul(ld(cf + 2 * jj), ld(avec + 2 * jj)) : ld(cf + 2 * jj), ld(g + 2 * jj), t);
            }
            for (int o = 4; o; o >>= 1) {
                t.re += __shfl_xor_sync(0xffffffffu, t.re, o);
                t.im += __shfl_xor_sync(0xffffffffu, t.im, o);
            }
            if (l == 0 && m < nz) sm[m] = mul(t, ld(chirp + 2 * m));
        }
    }
    fft_convolve(sm, fl, ftw, filt);
    const double sc = 1.0 / (double)fl;
    for (int j = threadIdx.x; j < nz; j += blockDim.x) {
        const int zr = nz - 1 - j;
        c2 out = mul(ld(bb + 2 * j), scal(sc, mul(sm[j], ld(chirp + 2 * j))));
        if (V1) {
            int64_t c0 = (h == 0) ? zr : zr + (int64_t)nz * (nx - 1);
            int64_t c1 = (h == 0) ? zr + nz : zr + (int64_t)nz * (nx - 2);
            out = fma2(cd1, ld(V1 + 2 * c0 * ldv), out);
            out = fma2(cd2, ld(V1 + 2 * c1 * ldv), out);
        }
        st(y + 2 * (int64_t)(h * nz + zr) * ldy, out);
    }
}

// u = C2T x for a vector on the interior grid (2 nz values)
__global__ void wep_c2t_kernel(int nx, int nz, double d1, double d2, const double* __restrict__ X, int64_t ldx, double* __restrict__ u) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= 2 * nz) return;
    int h = r / nz, z = r - h * nz;
    int64_t c0 = (h == 0) ? z : z + (int64_t)nz * (nx - 1);
    int64_t c1 = (h == 0) ? z + nz : z + (int64_t)nz * (nx - 2);
    st(u + 2 * r, add(scal(d1, ld(X + 2 * c0 * ldx)), scal(d2, ld(X + 2 * c1 * ldx))));
}

inline c2 cmul(c2 a, c2 b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

int boundary(const nepb_wep* h, int na, const double* d_coef, int ldc, const double* d_avec, const double* d_v2, int64_t ldv2, const double* d_V1, int64_t ldv, c2 cd1,
             c2 cd2, double* d_y, int64_t ldy) {
    static const bool direct = getenv("NEPB_WEP_DIRECT_DFT") && atoi(getenv("NEPB_WEP_DIRECT_DFT")) != 0;
    if (h->fl && !direct) {
        NEPB_CUDA(h->G.reserve((size_t)4 * h->nz * na));
        const size_t smem = (size_t)h->fl * 32;  // data + per-stage twiddle runs
        NEPB_LAUNCH(wep_czt_inv_kernel, dim3(na, 2), 512, smem, h->nz, na, h->fl, h->ftw.p, h->chirp.p, h->filt_i.p, h->bb.p, d_v2, ldv2, h->G.p);
        const double* tvec = nullptr;
        if (na > 4) {
            NEPB_CUDA(h->t.reserve((size_t)4 * h->nz));
            NEPB_LAUNCH(wep_combine_kernel, (2 * h->nz + 7) / 8, 256, 0, h->nz, na, h->chirp.p, d_coef, ldc, d_avec, h->G.p, h->t.p);
            tvec = h->t.p;
        }
        NEPB_LAUNCH(wep_czt_fwd_kernel, 2, 512, smem, h->nx, h->nz, na, h->fl, h->ftw.p, h->chirp.p, h->filt_f.p, h->bb.p, d_coef, ldc, d_avec,
                    h->G.p, tvec, d_V1, ldv, cd1, cd2, d_y, ldy);
        NEPB_LAUNCH_CHECK();
        return NEPB_OK;
    }
    NEPB_CUDA(h->t.reserve((size_t)4 * h->nz));
    const int grid = 2 * ((h->nz + WB_M - 1) / WB_M);
    NEPB_LAUNCH(wep_boundary_inv_kernel, grid, WB_M * WB_S, 0, h->nz, na, h->tw.p, h->bb.p, d_coef, ldc, d_avec, d_v2, ldv2, h->t.p);
    NEPB_LAUNCH(wep_boundary_fwd_kernel, grid, WB_M * WB_S, 0, h->nx, h->nz, h->tw.p, h->bb.p, h->t.p, d_V1, ldv, cd1, cd2, d_y, ldy);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

}  // namespace

extern "C" {

int nepb_wep_create(int nx, int nz, double hx, double hz, const double* K_scaled, const double* k_bar, const double* bb, nepb_wep** out) {
    NEPB_CHECK_ARG(out && K_scaled && k_bar && bb && nx >= 3 && nz >= 1 && hx > 0 && hz > 0, "bad arguments");
    NEPB_CHECK_ARG((int64_t)nz * nz < ((int64_t)1 << 62), "nz too large");
    nepb_wep* h = new nepb_wep();
    h->nx = nx;
    h->nz = nz;
    h->hx = hx;
    h->hz = hz;
    h->kbar_re = k_bar[0];
    h->kbar_im = k_bar[1];
    std::vector<double> tw(2 * (size_t)nz);
    for (int j = 0; j < nz; ++j) {  // exp(2 pi i j / nz), argument reduced to [-pi/4, pi/4] octants by the library
        long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)nz;
        tw[2 * j] = (double)cosl(ang);
        tw[2 * j + 1] = (double)sinl(ang);
    }
    // chirp-z data (FFT length up to 4096 = 128 KB of shared memory with the twiddles, i.e. nz <= 2048; longer boundaries keep the direct transform)
    std::vector<double> ftw, chirp, filt_f, filt_i;
    {
        int fl = 1, flog = 0;
        while (fl < 2 * nz - 1) fl <<= 1, ++flog;
        if (fl <= 4096) {
            typedef long double ld_t;
            const ld_t pi = 3.14159265358979323846264338327950288L;
            h->fl = fl;
            h->flog = flog;
            ftw.assign(2 * (size_t)fl, 0.0);  // per-stage runs: entry half + pos = exp(-i pi pos / half), half = 1, 2, ..., fl / 2
            for (int half = 1; half < fl; half <<= 1)
                for (int pos = 0; pos < half; ++pos) {
                    ftw[2 * (half + pos)] = (double)cosl(-pi * pos / half);
                    ftw[2 * (half + pos) + 1] = (double)sinl(-pi * pos / half);
                }
            chirp.resize(2 * (size_t)nz);
            std::vector<ld_t> br(fl, 0), bi(fl, 0);
            for (int k = 0; k < nz; ++k) {
                const long long k2 = ((long long)k * k) % (2LL * nz);  // exact: exp(-i pi k^2 / nz) has period 2 nz in k^2
                const ld_t ang = pi * (ld_t)k2 / (ld_t)nz;
                chirp[2 * k] = (double)cosl(ang);
                chirp[2 * k + 1] = (double)(-sinl(ang));
                // forward filter b[j] = conj(chirp)[|j|] = exp(+i pi j^2 / nz) at the indices j mod fl, -nz < j < nz
                br[k] = cosl(ang);
                bi[k] = sinl(ang);
                if (k) {
                    br[fl - k] = cosl(ang);
                    bi[fl - k] = sinl(ang);
                }
            }
            // extended-precision radix-2 FFT of the filter (decimation in time on a bit-reversed copy)
            std::vector<ld_t> xr(fl), xi(fl);
            auto brev = [&](int i) {
                int r = 0;
                for (int b = 0; b < flog; ++b) r |= ((i >> b) & 1) << (flog - 1 - b);
                return r;
            };
            for (int i = 0; i < fl; ++i) {
                xr[brev(i)] = br[i];
                xi[brev(i)] = bi[i];
            }
            for (int half = 1; half < fl; half <<= 1)
                for (int g0 = 0; g0 < fl; g0 += 2 * half)
                    for (int pos = 0; pos < half; ++pos) {
                        const ld_t ang = -pi * pos / half;
                        const ld_t wr = cosl(ang), wi = sinl(ang);
                        const int i0 = g0 + pos, i1 = i0 + half;
                        const ld_t vr = xr[i1] * wr - xi[i1] * wi, vi = xr[i1] * wi + xi[i1] * wr;
                        xr[i1] = xr[i0] - vr;
                        xi[i1] = xi[i0] - vi;
                        xr[i0] += vr;
                        xi[i0] += vi;
                    }
            filt_f.resize(2 * (size_t)fl);
            filt_i.resize(2 * (size_t)fl);
            for (int i = 0; i < fl; ++i) {
                const int r = brev(i);
                filt_f[2 * i] = (double)xr[r];
                filt_f[2 * i + 1] = (double)xi[r];
                const int rn = (fl - r) % fl;  // spectrum of the conjugate filter: conj(B[-r])
                filt_i[2 * i] = (double)xr[rn];
                filt_i[2 * i + 1] = (double)(-xi[rn]);
            }
        }
    }
    cudaError_t e = h->K.alloc((size_t)2 * nx * nz);
    if (h->fl) {
        auto up = [&](DevBuf<double>& b, const std::vector<double>& v) {
            if (e == cudaSuccess) e = b.alloc(v.size());
            if (e == cudaSuccess) e = cudaMemcpy(b.p, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice);
        };
        up(h->ftw, ftw);
        up(h->chirp, chirp);
        up(h->filt_f, filt_f);
        up(h->filt_i, filt_i);
        if (e == cudaSuccess && (size_t)h->fl * 32 > 48 * 1024) {
            e = cudaFuncSetAttribute(wep_czt_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->fl * 32);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(wep_czt_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->fl * 32);
        }
    }
    if (e == cudaSuccess) e = h->tw.alloc(2 * (size_t)nz);
    if (e == cudaSuccess) e = h->bb.alloc(2 * (size_t)nz);
    if (e == cudaSuccess) e = cudaMemcpy(h->K.p, K_scaled, sizeof(double) * 2 * nx * nz, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->tw.p, tw.data(), sizeof(double) * 2 * nz, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->bb.p, bb, sizeof(double) * 2 * nz, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("nepb_wep_create: %s", cudaGetErrorString(e));
        delete h;
        return e == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;
    }
    *out = h;
    return NEPB_OK;
}

int nepb_wep_destroy(nepb_wep* h) {
    delete h;
    return NEPB_OK;
}

// Derivative table of the boundary functions at one lambda: D (2 nz x ncols complex, row-major), D[m, j] = 1im * d^j/dlambda^j
// sqrt(beta_m(lambda)) (+ d0 for j = 0), Waveguide.jl:351-361.  Kept in HBM; nepb_wep_mlincomb_block with coef = NULL uses it with
// the coefficients a_j applied on the device, so a solver loop at a fixed shift uploads 16 na bytes per call.
int nepb_wep_set_table(nepb_wep* h, int ncols, const double* D) {
    NEPB_CHECK_ARG(h && D && ncols >= 1, "bad arguments");
    NEPB_CUDA(cudaStreamSynchronize(stream()));  // calls in flight may still read the old table
    NEPB_CUDA(h->table.reserve((size_t)4 * h->nz * ncols));
    NEPB_CUDA(cudaMemcpy(h->table.p, D, sizeof(double) * 4 * h->nz * ncols, cudaMemcpyHostToDevice));
    h->table_cols = ncols;
    return NEPB_OK;
}

int nepb_wep_info(const nepb_wep* h, int* nx, int* nz, int64_t* n) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    if (nx) *nx = h->nx;
    if (nz) *nz = h->nz;
    if (n) *n = (int64_t)h->nx * h->nz + 2 * h->nz;
    return NEPB_OK;
}

int nepb_wep_mlincomb_block(const nepb_wep* h, const double* lambda, const nepb_block* V, int vcol0, int na, const double* a,
                            const double* coef, nepb_block* Z, int zcol) {
    NEPB_CHECK_ARG(h && lambda && V && Z && a, "NULL argument");
    NEPB_CHECK_ARG(coef || h->table_cols >= na, "no coefficient block given and the table set by nepb_wep_set_table has %d < %d columns",
                   h->table_cols, na);
    int64_t m = (int64_t)h->nx * h->nz, n = m + 2 * h->nz;
    NEPB_CHECK_ARG(V->n == n && Z->n == n, "Incompatible sizes: Length of vectors = %lld, size of NEP = %lld.", (long long)V->n, (long long)n);
    NEPB_CHECK_ARG(na >= 1 && vcol0 >= 0 && vcol0 + na <= V->k && zcol >= 0 && zcol < Z->k, "column range outside the block");
    NEPB_CHECK_ARG(V != Z, "V and Z must be different blocks");
    const double* d_coef = h->table.p;
    const double* d_avec = nullptr;
    int ldc = h->table_cols;
    if (coef) {
        NEPB_CUDA(h->coef.reserve((size_t)4 * h->nz * na));
        NEPB_CUDA(cudaMemcpyAsync(h->coef.p, coef, sizeof(double) * 4 * h->nz * na, cudaMemcpyHostToDevice, stream()));
        d_coef = h->coef.p;
        ldc = na;
    } else {
        NEPB_CUDA(h->avec.reserve((size_t)2 * na));
        NEPB_CUDA(cudaMemcpyAsync(h->avec.p, a, sizeof(double) * 2 * na, cudaMemcpyHostToDevice, stream()));
        d_avec = h->avec.p;
    }
    InteriorArgs p;
    p.nx = h->nx;
    p.nz = h->nz;
    p.na = na < 3 ? na : 3;
    p.ihx2 = 1.0 / (h->hx * h->hx);
    p.ihz2 = 1.0 / (h->hz * h->hz);
    p.ihz = 1.0 / (2.0 * h->hz);
    p.lam = {lambda[0], lambda[1]};
    c2 l2 = cmul(p.lam, p.lam);
    p.lam2k = {l2.re + h->kbar_re, l2.im + h->kbar_im};
    p.a0 = {a[0], a[1]};
    p.a1 = na > 1 ? c2{a[2], a[3]} : c2{0, 0};
    p.a2 = na > 2 ? c2{a[4], a[5]} : c2{0, 0};
    p.c1s = {a[0] * p.ihx2, a[1] * p.ihx2};
    const double* Vp = V->d.p + 2 * (int64_t)vcol0;
    double* Zp = Z->d.p + 2 * (int64_t)zcol;
    const double* v2 = Vp + 2 * m * V->k;
    NEPB_LAUNCH(wep_interior_kernel, (unsigned)((m + 255) / 256), 256, 0, p, h->K.p, Vp, (int64_t)V->k, v2, (int64_t)V->k, Zp, (int64_t)Z->k);
    double d1 = 2.0 / h->hx, d2 = -1.0 / (2.0 * h->hx);
    return boundary(h, na, d_coef, ldc, d_avec, v2, V->k, Vp, V->k, c2{a[0] * d1, a[1] * d1}, c2{a[0] * d2, a[1] * d2}, Zp + 2 * m * Z->k, Z->k);
}

// x, y: 2 nz host vectors; coef: 2 nz complex (1 ./ [sM; sP] for the reference's Pinv)
int nepb_wep_pinv(const nepb_wep* h, const double* coef, const double* x, double* y) {
    NEPB_CHECK_ARG(h && coef && x && y, "NULL argument");
    size_t len = (size_t)4 * h->nz;
    NEPB_CUDA(h->coef.reserve(len));
    NEPB_CUDA(h->u.reserve(len));
    NEPB_CUDA(h->w.reserve(len));
    NEPB_CUDA(cudaMemcpyAsync(h->coef.p, coef, sizeof(double) * len, cudaMemcpyHostToDevice, stream()));
    NEPB_CUDA(cudaMemcpyAsync(h->u.p, x, sizeof(double) * len, cudaMemcpyHostToDevice, stream()));
    int rc = boundary(h, 1, h->coef.p, 1, nullptr, h->u.p, 1, nullptr, 0, c2{0, 0}, c2{0, 0}, h->w.p, 1);
    if (rc != NEPB_OK) return rc;
    NEPB_CUDA(cudaMemcpyAsync(y, h->w.p, sizeof(double) * len, cudaMemcpyDeviceToHost, stream()));
    NEPB_CUDA(cudaStreamSynchronize(stream()));
    return NEPB_OK;
}

// Y[:, ycol] = SchurMatVec(lambda) * X[:, xcol]: blocks with nx*nz rows; sinv = 1 ./ [sM(lambda); sP(lambda)] (2 nz complex, host)
int nepb_wep_schur_matvec_block(const nepb_wep* h, const double* lambda, const double* sinv, const nepb_block* X, int xcol, nepb_block* Y,
                                int ycol) {
    NEPB_CHECK_ARG(h && lambda && sinv && X && Y && X != Y, "bad arguments");
    int64_t m = (int64_t)h->nx * h->nz;
    NEPB_CHECK_ARG(X->n == m && Y->n == m && xcol >= 0 && xcol < X->k && ycol >= 0 && ycol < Y->k, "blocks must have nx*nz rows");
    size_t len = (size_t)4 * h->nz;
    NEPB_CUDA(h->coef.reserve(len));
    NEPB_CUDA(h->u.reserve(len));
    NEPB_CUDA(h->w.reserve(len));
    NEPB_CUDA(cudaMemcpyAsync(h->coef.p, sinv, sizeof(double) * len, cudaMemcpyHostToDevice, stream()));
    const double* Xp = X->d.p + 2 * (int64_t)xcol;
    NEPB_LAUNCH(wep_c2t_kernel, (2 * h->nz + 127) / 128, 128, 0, h->nx, h->nz, 2.0 / h->hx, -1.0 / (2.0 * h->hx), Xp, (int64_t)X->k, h->u.p);
    int rc = boundary(h, 1, h->coef.p, 1, nullptr, h->u.p, 1, nullptr, 0, c2{0, 0}, c2{0, 0}, h->w.p, 1);
    if (rc != NEPB_OK) return rc;
    InteriorArgs p;
    p.nx = h->nx;
    p.nz = h->nz;
    p.na = 1;
    p.ihx2 = 1.0 / (h->hx * h->hx);
    p.ihz2 = 1.0 / (h->hz * h->hz);
    p.ihz = 1.0 / (2.0 * h->hz);
    p.lam = {lambda[0], lambda[1]};
    c2 l2 = cmul(p.lam, p.lam);
    p.lam2k = {l2.re + h->kbar_re, l2.im + h->kbar_im};
    p.a0 = {1, 0};
    p.a1 = p.a2 = {0, 0};
    p.c1s = {-p.ihx2, 0};
    NEPB_LAUNCH(wep_interior_kernel, (unsigned)((m + 255) / 256), 256, 0, p, h->K.p, Xp, (int64_t)X->k, h->w.p, (int64_t)1,
                Y->d.p + 2 * (int64_t)ycol, (int64_t)Y->k);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

// algorithmic HBM bytes of one nepb_wep_mlincomb_block call (bench roofline): X, K and the result once per grid point,
// the used derivative columns, the boundary rows of all na columns
int64_t nepb_wep_mlincomb_bytes(const nepb_wep* h, int na) {
    if (!h) return 0;
    int64_t m = (int64_t)h->nx * h->nz;
    int nd = na < 3 ? na : 3;
    return m * 16 * (2 + nd) + (int64_t)2 * h->nz * 16 * (na + 1);
}

}  // extern "C"
