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
// O(nz^2) transform with an exact twiddle table (index m k mod nz in integer arithmetic): for nz = 945 that is 1.8 M
// complex multiply-adds per transform spread over 2 nz CTAs -- microseconds, no FFT plan, no library, and with the
// tree-shaped CTA sums an error of a few ulp.  Sums have a fixed order: results are bitwise reproducible.
#include "common.h"

#include <cmath>

using namespace nepb;

struct nepb_wep {
    int nx = 0, nz = 0;
    double hx = 0, hz = 0;
    double kbar_re = 0, kbar_im = 0;
    DevBuf<double> K;    // nz*nx complex, K - k_bar, index z + nz*x (the reference's vec(K))
    DevBuf<double> tw;   // nz complex: exp(2 pi i j / nz)
    DevBuf<double> bb;   // nz complex: the reference's bb (|bb| = 1, bbinv = conj(bb))
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

__device__ __forceinline__ c2 block_sum(c2 v, c2* sh) {
    for (int o = 16; o; o >>= 1) {
        v.re += __shfl_down_sync(0xffffffffu, v.re, o);
        v.im += __shfl_down_sync(0xffffffffu, v.im, o);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    c2 r{0, 0};
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r = add(r, sh[i]);
    return r;  // valid in thread 0
}

// t[m] = (1/nz) sum_k W^{+mm k} conj(bb[k]) * ( sum_jj coef[m, jj] v2[h nz + nz-1-k, jj] ),  m = h nz + mm
// = the reference's  sum_jj coef[:, jj] .* [Rinv(v2_jj[1:nz]); Rinv(v2_jj[nz+1:2nz])]   (Waveguide.jl:366-373, Rinv :169-171)
__global__ void __launch_bounds__(128) wep_boundary_inv_kernel(int nz, int na, const double* __restrict__ tw,
                                                               const double* __restrict__ bb, const double* __restrict__ coef,
                                                               const double* __restrict__ v2, int64_t ldv2, double* __restrict__ t) {
    __shared__ c2 sh[4];
    int m = blockIdx.x, h = m / nz, mm = m - h * nz;
    const double* cf = coef + 2 * (int64_t)m * na;
    c2 acc{0, 0};
    for (int k = threadIdx.x; k < nz; k += blockDim.x) {
        const double* row = v2 + 2 * (int64_t)(h * nz + nz - 1 - k) * ldv2;
        c2 s{0, 0};
        for (int jj = 0; jj < na; ++jj) s = fma2(ld(cf + 2 * jj), ld(row + 2 * jj), s);
        int idx = (int)(((int64_t)mm * k) % nz);
        c2 b = ld(bb + 2 * k);
        c2 f = mul(ld(tw + 2 * idx), c2{b.re, -b.im});
        acc = fma2(f, s, acc);
    }
    c2 r = block_sum(acc, sh);
    if (threadIdx.x == 0) st(t + 2 * m, scal(1.0 / nz, r));
}

// y[h nz + nz-1-j] = bb[j] sum_m t[h nz + m] W^{-j m}  (+ the C2T rows: d1 * first / last grid column + d2 * its neighbour)
// = R(t[1:nz]), R(t[nz+1:2nz])   (Waveguide.jl:374-376, R :165-167; C2T waveguide_FD.jl:52-60)
__global__ void __launch_bounds__(128) wep_boundary_fwd_kernel(int nx, int nz, const double* __restrict__ tw,
                                                               const double* __restrict__ bb, const double* __restrict__ t,
                                                               const double* __restrict__ V1, int64_t ldv, c2 cd1, c2 cd2,
                                                               double* __restrict__ y, int64_t ldy) {
    __shared__ c2 sh[4];
    int j = blockIdx.x % nz, h = blockIdx.x / nz;
    c2 acc{0, 0};
    for (int m = threadIdx.x; m < nz; m += blockDim.x) {
        int idx = (int)(((int64_t)j * m) % nz);
        c2 w = ld(tw + 2 * idx);
        acc = fma2(c2{w.re, -w.im}, ld(t + 2 * (h * nz + m)), acc);
    }
    c2 r = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        int zr = nz - 1 - j;
        c2 out = mul(ld(bb + 2 * j), r);
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

int boundary(const nepb_wep* h, int na, const double* d_coef, const double* d_v2, int64_t ldv2, const double* d_V1, int64_t ldv, c2 cd1,
             c2 cd2, double* d_y, int64_t ldy) {
    NEPB_CUDA(h->t.reserve((size_t)4 * h->nz));
    NEPB_LAUNCH(wep_boundary_inv_kernel, 2 * h->nz, 128, 0, h->nz, na, h->tw.p, h->bb.p, d_coef, d_v2, ldv2, h->t.p);
    NEPB_LAUNCH(wep_boundary_fwd_kernel, 2 * h->nz, 128, 0, h->nx, h->nz, h->tw.p, h->bb.p, h->t.p, d_V1, ldv, cd1, cd2, d_y, ldy);
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
    cudaError_t e = h->K.alloc((size_t)2 * nx * nz);
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

int nepb_wep_info(const nepb_wep* h, int* nx, int* nz, int64_t* n) {
    NEPB_CHECK_ARG(h, "handle is NULL");
    if (nx) *nx = h->nx;
    if (nz) *nz = h->nz;
    if (n) *n = (int64_t)h->nx * h->nz + 2 * h->nz;
    return NEPB_OK;
}

int nepb_wep_mlincomb_block(const nepb_wep* h, const double* lambda, const nepb_block* V, int vcol0, int na, const double* a,
                            const double* coef, nepb_block* Z, int zcol) {
    NEPB_CHECK_ARG(h && lambda && V && Z && a && coef, "NULL argument");
    int64_t m = (int64_t)h->nx * h->nz, n = m + 2 * h->nz;
    NEPB_CHECK_ARG(V->n == n && Z->n == n, "Incompatible sizes: Length of vectors = %lld, size of NEP = %lld.", (long long)V->n, (long long)n);
    NEPB_CHECK_ARG(na >= 1 && vcol0 >= 0 && vcol0 + na <= V->k && zcol >= 0 && zcol < Z->k, "column range outside the block");
    NEPB_CHECK_ARG(V != Z, "V and Z must be different blocks");
    NEPB_CUDA(h->coef.reserve((size_t)4 * h->nz * na));
    NEPB_CUDA(cudaMemcpyAsync(h->coef.p, coef, sizeof(double) * 4 * h->nz * na, cudaMemcpyHostToDevice, stream()));
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
    return boundary(h, na, h->coef.p, v2, V->k, Vp, V->k, c2{a[0] * d1, a[1] * d1}, c2{a[0] * d2, a[1] * d2}, Zp + 2 * m * Z->k, Z->k);
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
    int rc = boundary(h, 1, h->coef.p, h->u.p, 1, nullptr, 0, c2{0, 0}, c2{0, 0}, h->w.p, 1);
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
    int rc = boundary(h, 1, h->coef.p, h->u.p, 1, nullptr, 0, c2{0, 0}, c2{0, 0}, h->w.p, 1);
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
