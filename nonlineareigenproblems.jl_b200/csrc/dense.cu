// Dense tall-skinny blocks of the infinite-Arnoldi callers, resident in HBM (sm_100a).
//
// Replaces (reference, relative to src/): the Gram-Schmidt call orthogonalize_and_normalize!(V, w, h, DGKS()) of
// method_iar.jl:107, method_tiar.jl:128, method_nleigs.jl:293 (IterativeSolvers 0.9.2: classical Gram-Schmidt with
// DGKS re-orthogonalisation), and the tall-skinny products Z*a' (method_tiar.jl:119,188), VV*W (:189), Q = VV*Z
// (method_iar.jl:115).
//
//   orth   : two bandwidth-bound passes over the basis per Gram-Schmidt sweep (h = V^H w, then w -= V h fused with the
//            norm), partial sums reduced in a fixed order -> bitwise reproducible.
//   gemm   : Y = A * C with A (rows x ka), C (ka x q) small: FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64 -- tcgen05
//            has no FP64 kind), complex product as four real MMAs on split re/im planes staged in shared memory.
// Blocks are row-major n x k complex (nepb_block), so a row of the basis is contiguous.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.h"

namespace nepb {

// ---------------------------------------------------------------------------------------------
// h_partial[cta][j] = sum_{r in slab} conj(V[r, j]) * w[r]
// ---------------------------------------------------------------------------------------------
constexpr int ORTH_MAXK = 256;  // basis columns handled per call
constexpr int ORTH_CPL = ORTH_MAXK / 32;

__global__ void __launch_bounds__(256) orth_dot_kernel(int64_t rows, int k, const double2* __restrict__ V, int ldv, const double2* __restrict__ w,
                                                       int ldw, double2* __restrict__ partial) {
    __shared__ double2 red[8][ORTH_MAXK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 acc[ORTH_CPL];
#pragma unroll
    for (int c = 0; c < ORTH_CPL; ++c) acc[c] = make_double2(0.0, 0.0);
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    for (int64_t r = r0 + warp; r < r1; r += 8) {
        const double2 wr = w[(size_t)r * ldw];
        const double2* vr = V + (size_t)r * ldv;
#pragma unroll
        for (int c = 0; c < ORTH_CPL; ++c) {
            const int j = lane + 32 * c;
            if (j < k) {
                const double2 v = vr[j];
                // conj(v) * w
                acc[c].x = fma(v.x, wr.x, acc[c].x);
                acc[c].x = fma(v.y, wr.y, acc[c].x);
                acc[c].y = fma(v.x, wr.y, acc[c].y);
                acc[c].y = fma(-v.y, wr.x, acc[c].y);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < ORTH_CPL; ++c) red[warp][lane + 32 * c] = acc[c];
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += 256) {
        double2 s = red[0][j];
        for (int q = 1; q < 8; ++q) {
            s.x += red[q][j].x;
            s.y += red[q][j].y;
        }
        partial[(size_t)blockIdx.x * k + j] = s;
    }
}

// out[j] (+)= sum_cta partial[cta][j] in cta order; accumulate != 0 adds to the existing value (DGKS: h .+= correction)
__global__ void __launch_bounds__(256) orth_reduce_kernel(int ncta, int k, const double2* __restrict__ partial, double2* __restrict__ hcur,
                                                          double2* __restrict__ hsum, int accumulate, double* __restrict__ hnorm2) {
    __shared__ double sn[256];
    double nn = 0.0;
    for (int j = threadIdx.x; j < k; j += 256) {
        double2 s = make_double2(0.0, 0.0);
        for (int c = 0; c < ncta; ++c) {
            s.x += partial[(size_t)c * k + j].x;
            s.y += partial[(size_t)c * k + j].y;
        }
        hcur[j] = s;
        if (accumulate) {
            hsum[j].x += s.x;
            hsum[j].y += s.y;
        } else {
            hsum[j] = s;
        }
        nn += s.x * s.x + s.y * s.y;
    }
    sn[threadIdx.x] = nn;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 256; ++i) t += sn[i];
        *hnorm2 = t;  // ||h||^2 of this sweep's projection
    }
}

// w[r] -= sum_j V[r, j] h[j]; npartial[cta] = sum |w[r]|^2 over the slab
__global__ void __launch_bounds__(256) orth_update_kernel(int64_t rows, int k, const double2* __restrict__ V, int ldv, double2* __restrict__ w,
                                                          int ldw, const double2* __restrict__ h, double* __restrict__ npartial) {
    __shared__ double2 sh[ORTH_MAXK];
    __shared__ double sn[8];
    for (int j = threadIdx.x; j < k; j += 256) sh[j] = h[j];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    double nn = 0.0;
    for (int64_t r = r0 + warp; r < r1; r += 8) {
        const double2* vr = V + (size_t)r * ldv;
        double2 s = make_double2(0.0, 0.0);
        for (int j = lane; j < k; j += 32) {
            const double2 v = vr[j], hj = sh[j];
            s.x = fma(v.x, hj.x, s.x);
            s.x = fma(-v.y, hj.y, s.x);
            s.y = fma(v.x, hj.y, s.y);
            s.y = fma(v.y, hj.x, s.y);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s.x += __shfl_xor_sync(0xffffffffu, s.x, off);
            s.y += __shfl_xor_sync(0xffffffffu, s.y, off);
        }
        if (lane == 0) {
            double2 wr = w[(size_t)r * ldw];
            wr.x -= s.x;
            wr.y -= s.y;
            w[(size_t)r * ldw] = wr;
            nn += wr.x * wr.x + wr.y * wr.y;
        }
    }
    if (lane == 0) sn[warp] = nn;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += sn[q];
        npartial[blockIdx.x] = t;
    }
}

__global__ void orth_norm_reduce_kernel(int ncta, const double* __restrict__ npartial, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int c = 0; c < ncta; ++c) t += npartial[c];
        *out = t;
    }
}

// x[r*ld] *= alpha (real), optional copy into a second location
__global__ void __launch_bounds__(256) scale_col_kernel(int64_t rows, double2* __restrict__ w, int ldw, double alpha) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double2 v = w[(size_t)r * ldw];
    v.x *= alpha;
    v.y *= alpha;
    w[(size_t)r * ldw] = v;
}

// dst[:, d0 + c] = alpha * src[:, s0 + c], c < nc   (row-major blocks)
__global__ void __launch_bounds__(256) copy_cols_kernel(int64_t rows, int nc, const double2* __restrict__ src, int lds, int s0,
                                                        double2* __restrict__ dst, int ldd, int d0, double2 alpha) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nc) return;
    const int64_t r = idx / nc;
    const int c = (int)(idx % nc);
    const double2 v = src[(size_t)r * lds + s0 + c];
    dst[(size_t)r * ldd + d0 + c] = make_double2(alpha.x * v.x - alpha.y * v.y, alpha.x * v.y + alpha.y * v.x);
}

// ---------------------------------------------------------------------------------------------
// Y[rows x q] = A[rows x ka] * C[ka x q]  with FP64 tensor cores.
// CTA: 4 warps, 32 rows x 64 columns; K in chunks of 32; operands as split re / im planes in shared memory.
// mma.sync.aligned.m8n8k4.row.col.f64: a0 = A[lane>>2][lane&3], b0 = B[lane&3][lane>>2], c{0,1} = C[lane>>2][2*(lane&3)+{0,1}]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int GM_R = 32, GM_C = 64, GM_K = 16;
__global__ void __launch_bounds__(128) block_gemm_dmma_kernel(int64_t rows, int ka, int q, const double2* __restrict__ A, int lda,
                                                              const double2* __restrict__ C /* row-major ka x q */, double2* __restrict__ Y,
                                                              int ldy) {
    __shared__ double sAr[GM_R][GM_K + 1], sAi[GM_R][GM_K + 1];
    __shared__ double sCr[GM_K][GM_C + 1], sCi[GM_K][GM_C + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * GM_R;
    const int col0 = blockIdx.y * GM_C;
    double cr[8][2], ci[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) cr[t][0] = cr[t][1] = ci[t][0] = ci[t][1] = 0.0;
    const int ar = lane >> 2, ak = lane & 3;
    for (int k0 = 0; k0 < ka; k0 += GM_K) {
        __syncthreads();
        for (int idx = tid; idx < GM_R * GM_K; idx += 128) {
            const int r = idx / GM_K, kk = idx % GM_K;
            double2 v = make_double2(0.0, 0.0);
            if (row0 + r < rows && k0 + kk < ka) v = A[(size_t)(row0 + r) * lda + k0 + kk];
            sAr[r][kk] = v.x;
            sAi[r][kk] = v.y;
        }
        for (int idx = tid; idx < GM_K * GM_C; idx += 128) {
            const int kk = idx / GM_C, c = idx % GM_C;
            double2 v = make_double2(0.0, 0.0);
            if (k0 + kk < ka && col0 + c < q) v = C[(size_t)(k0 + kk) * q + col0 + c];
            sCr[kk][c] = v.x;
            sCi[kk][c] = v.y;
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < GM_K; k4 += 4) {
            const double a_r = sAr[warp * 8 + ar][k4 + ak];
            const double a_i = sAi[warp * 8 + ar][k4 + ak];
            const double na_i = -a_i;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const double b_r = sCr[k4 + ak][t * 8 + ar];
                const double b_i = sCi[k4 + ak][t * 8 + ar];
                dmma(cr[t][0], cr[t][1], a_r, b_r);
                dmma(cr[t][0], cr[t][1], na_i, b_i);
                dmma(ci[t][0], ci[t][1], a_r, b_i);
                dmma(ci[t][0], ci[t][1], a_i, b_r);
            }
        }
    }
    const int64_t r = row0 + warp * 8 + ar;
    if (r < rows) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int c = col0 + t * 8 + 2 * ak;
            if (c < q) Y[(size_t)r * ldy + c] = make_double2(cr[t][0], ci[t][0]);
            if (c + 1 < q) Y[(size_t)r * ldy + c + 1] = make_double2(cr[t][1], ci[t][1]);
        }
    }
}

// iar's block shift (method_iar.jl:100-101): Y[i, ycol0 + b] = V[b*n + i, vcol] / (b + 1), b < nb
__global__ void __launch_bounds__(256) iar_expand_kernel(int64_t n, int nb, const double2* __restrict__ V, int ldv, int vcol, double2* __restrict__ Y,
                                                         int ldy, int ycol0, int scale) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * nb) return;
    const int64_t i = idx / nb;
    const int b = (int)(idx % nb);
    double2 v = V[((size_t)b * n + i) * ldv + vcol];
    if (scale) {
        const double sc = 1.0 / (double)(b + 1);
        v.x *= sc;
        v.y *= sc;
    }
    Y[(size_t)i * ldy + ycol0 + b] = v;
}
// vv = vec(y[:, 0:nb)) (method_iar.jl:105): V[b*n + i, vcol] = Y[i, ycol0 + b]
__global__ void __launch_bounds__(256) iar_pack_kernel(int64_t n, int nb, const double2* __restrict__ Y, int ldy, int ycol0, double2* __restrict__ V,
                                                       int ldv, int vcol) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * nb) return;
    const int64_t i = idx / nb;
    const int b = (int)(idx % nb);
    V[((size_t)b * n + i) * ldv + vcol] = Y[(size_t)i * ldy + ycol0 + b];
}
// out[c] = sum_r |A[r, c0 + c]|^2 partial per CTA (column norms of Ritz residual blocks)
__global__ void __launch_bounds__(256) colnorm2_kernel(int64_t rows, int nc, const double2* __restrict__ A, int lda, int c0, double* __restrict__ partial) {
    __shared__ double red[8][ORTH_MAXK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[ORTH_CPL];
#pragma unroll
    for (int c = 0; c < ORTH_CPL; ++c) acc[c] = 0.0;
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    for (int64_t r = r0 + warp; r < r1; r += 8) {
#pragma unroll
        for (int c = 0; c < ORTH_CPL; ++c) {
            const int j = lane + 32 * c;
            if (j < nc) {
                const double2 v = A[(size_t)r * lda + c0 + j];
                acc[c] = fma(v.x, v.x, acc[c]);
                acc[c] = fma(v.y, v.y, acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < ORTH_CPL; ++c) red[warp][lane + 32 * c] = acc[c];
    __syncthreads();
    for (int j = threadIdx.x; j < nc; j += 256) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += red[q][j];
        partial[(size_t)blockIdx.x * nc + j] = t;
    }
}

struct OrthScratch {
    DevBuf<double2> partial, hcur, hsum;
    DevBuf<double> npartial, scal;
};
static OrthScratch g_orth;
static DevBuf<double> g_gemm_c;

}  // namespace nepb

using namespace nepb;

extern "C" {

// orthogonalize_and_normalize!(V[:, 0:k), w = W[:, wcol], h, DGKS): returns ||w|| before normalisation in *nrm_out,
// h[k] (host, complex) = accumulated projection coefficients, *sweeps = number of Gram-Schmidt sweeps (>= 1).
// V and W may be the same block (w is then a later column of the basis, method_tiar.jl:128).
int nepb_orth_dgks(const nepb_block* V, int k, nepb_block* W, int wcol, int64_t rows, double* h, double* nrm_out, int* sweeps) {
    NEPB_CHECK_ARG(V && W && h && nrm_out, "NULL argument");
    NEPB_CHECK_ARG(k >= 0 && k <= V->k && k <= ORTH_MAXK, "k=%d basis columns (block has %d, limit %d)", k, V->k, ORTH_MAXK);
    NEPB_CHECK_ARG(wcol >= 0 && wcol < W->k, "column %d out of range", wcol);
    if (rows <= 0) rows = V->n;
    NEPB_CHECK_ARG(rows <= V->n && rows <= W->n, "rows=%lld exceed the block height", (long long)rows);
    NEPB_CHECK_ARG(!(V == W && wcol < k), "w must not be one of the basis columns");
    const int ncta = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)sm_count() * 4);
    const int kk = std::max(k, 1);
    NEPB_CUDA(g_orth.partial.reserve((size_t)ncta * kk));
    NEPB_CUDA(g_orth.hcur.reserve(kk));
    NEPB_CUDA(g_orth.hsum.reserve(kk));
    NEPB_CUDA(g_orth.npartial.reserve(ncta));
    NEPB_CUDA(g_orth.scal.reserve(4));
    const double2* Vp = (const double2*)V->d.p;
    double2* wp = (double2*)W->d.p + wcol;
    const int ldv = V->k, ldw = W->k;
    double hs[2] = {0.0, 0.0};  // {||h||^2 of the last sweep, ||w||^2}
    int nsweep = 0;
    const double eta = 1.0 / std::sqrt(2.0);
    for (;;) {
        if (k > 0) {
            NEPB_LAUNCH(orth_dot_kernel, ncta, 256, 0, rows, k, Vp, ldv, (const double2*)wp, ldw, g_orth.partial.p);
            NEPB_LAUNCH(orth_reduce_kernel, 1, 256, 0, ncta, k, (const double2*)g_orth.partial.p, g_orth.hcur.p, g_orth.hsum.p, nsweep > 0 ? 1 : 0,
                        g_orth.scal.p);
        } else {
            NEPB_CUDA(cudaMemsetAsync(g_orth.scal.p, 0, sizeof(double), stream()));
        }
        NEPB_LAUNCH(orth_update_kernel, ncta, 256, 0, rows, k, Vp, ldv, wp, ldw, (const double2*)g_orth.hcur.p, g_orth.npartial.p);
        NEPB_LAUNCH(orth_norm_reduce_kernel, 1, 32, 0, ncta, (const double*)g_orth.npartial.p, g_orth.scal.p + 1);
        NEPB_LAUNCH_CHECK();
        NEPB_CUDA(cudaMemcpyAsync(hs, g_orth.scal.p, sizeof(hs), cudaMemcpyDeviceToHost, stream()));
        NEPB_CUDA(cudaStreamSynchronize(stream()));
        ++nsweep;
        const double nrm = std::sqrt(hs[1]), proj = std::sqrt(hs[0]);
        if (!(nrm < eta * proj) || nsweep >= 8 || k == 0) break;  // DGKS criterion (IterativeSolvers orthogonalize.jl)
    }
    const double nrm = std::sqrt(hs[1]);
    if (k > 0) NEPB_CUDA(cudaMemcpyAsync(h, g_orth.hsum.p, sizeof(double) * 2 * k, cudaMemcpyDeviceToHost, stream()));
    NEPB_LAUNCH(scale_col_kernel, (unsigned)((rows + 255) / 256), 256, 0, rows, wp, ldw, nrm > 0 ? 1.0 / nrm : 0.0);
    NEPB_LAUNCH_CHECK();
    NEPB_CUDA(cudaStreamSynchronize(stream()));
    *nrm_out = nrm;
    if (sweeps) *sweeps = nsweep;
    return NEPB_OK;
}

// Y[:, ycol0 : ycol0+q) = A[:, acol0 : acol0+ka) * C, C host column-major ka x q (ldc).  A and Y must not overlap in
// the written columns.
int nepb_block_gemm(const nepb_block* A, int acol0, int ka, const double* C, int64_t ldc, int q, nepb_block* Y, int ycol0, int64_t rows) {
    NEPB_CHECK_ARG(A && Y && C, "NULL argument");
    NEPB_CHECK_ARG(ka >= 1 && q >= 1 && acol0 >= 0 && acol0 + ka <= A->k && ycol0 >= 0 && ycol0 + q <= Y->k && ldc >= ka, "bad column windows");
    if (rows <= 0) rows = A->n;
    NEPB_CHECK_ARG(rows <= A->n && rows <= Y->n, "rows exceed the block height");
    NEPB_CHECK_ARG(!(A == Y && ycol0 < acol0 + ka && acol0 < ycol0 + q), "input and output columns overlap");
    std::vector<double> cr((size_t)2 * ka * q);  // row-major ka x q
    for (int j = 0; j < q; ++j)
        for (int i = 0; i < ka; ++i) {
            cr[2 * ((size_t)i * q + j)] = C[2 * ((size_t)j * ldc + i)];
            cr[2 * ((size_t)i * q + j) + 1] = C[2 * ((size_t)j * ldc + i) + 1];
        }
    NEPB_CUDA(g_gemm_c.reserve(cr.size()));
    NEPB_CUDA(cudaMemcpyAsync(g_gemm_c.p, cr.data(), cr.size() * sizeof(double), cudaMemcpyHostToDevice, stream()));
    NEPB_CUDA(cudaStreamSynchronize(stream()));
    dim3 grid((unsigned)((rows + GM_R - 1) / GM_R), (unsigned)((q + GM_C - 1) / GM_C));
    NEPB_LAUNCH(block_gemm_dmma_kernel, grid, 128, 0, rows, ka, q, (const double2*)A->d.p + acol0, A->k, (const double2*)g_gemm_c.p,
                (double2*)Y->d.p + ycol0, Y->k);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

int nepb_iar_expand(const nepb_block* V, int vcol, int64_t n, int nb, nepb_block* Y, int ycol0, int scale_by_index) {
    NEPB_CHECK_ARG(V && Y && n >= 1 && nb >= 1 && vcol >= 0 && vcol < V->k && ycol0 >= 0 && ycol0 + nb <= Y->k, "bad arguments");
    NEPB_CHECK_ARG(n * nb <= V->n && n <= Y->n, "row ranges exceed the blocks");
    NEPB_LAUNCH(iar_expand_kernel, (unsigned)((n * nb + 255) / 256), 256, 0, n, nb, (const double2*)V->d.p, V->k, vcol, (double2*)Y->d.p, Y->k, ycol0,
                scale_by_index);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

int nepb_iar_pack(const nepb_block* Y, int ycol0, int nb, int64_t n, nepb_block* V, int vcol) {
    NEPB_CHECK_ARG(V && Y && n >= 1 && nb >= 1 && vcol >= 0 && vcol < V->k && ycol0 >= 0 && ycol0 + nb <= Y->k, "bad arguments");
    NEPB_CHECK_ARG(n * nb <= V->n && n <= Y->n, "row ranges exceed the blocks");
    NEPB_LAUNCH(iar_pack_kernel, (unsigned)((n * nb + 255) / 256), 256, 0, n, nb, (const double2*)Y->d.p, Y->k, ycol0, (double2*)V->d.p, V->k, vcol);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

// out[c] = || A[0:rows, c0 + c] ||_2, c < nc (host doubles)
int nepb_block_colnorms(const nepb_block* A, int c0, int nc, int64_t rows, double* out) {
    NEPB_CHECK_ARG(A && out && nc >= 1 && nc <= ORTH_MAXK && c0 >= 0 && c0 + nc <= A->k, "bad arguments");
    if (rows <= 0) rows = A->n;
    NEPB_CHECK_ARG(rows <= A->n, "rows exceed the block height");
    const int ncta = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)sm_count() * 4);
    static DevBuf<double> part;
    NEPB_CUDA(part.reserve((size_t)ncta * nc));
    NEPB_LAUNCH(colnorm2_kernel, ncta, 256, 0, rows, nc, (const double2*)A->d.p, A->k, c0, part.p);
    NEPB_LAUNCH_CHECK();
    std::vector<double> hp((size_t)ncta * nc);
    NEPB_CUDA(cudaMemcpyAsync(hp.data(), part.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, stream()));
    NEPB_CUDA(cudaStreamSynchronize(stream()));
    for (int c = 0; c < nc; ++c) {
        double t = 0.0;
        for (int q = 0; q < ncta; ++q) t += hp[(size_t)q * nc + c];
        out[c] = std::sqrt(t);
    }
    return NEPB_OK;
}

// dst[:, d0 : d0+nc) = alpha * src[:, s0 : s0+nc)
int nepb_block_copy_cols(const nepb_block* src, int s0, int nc, nepb_block* dst, int d0, const double* alpha, int64_t rows) {
    NEPB_CHECK_ARG(src && dst && nc >= 1 && s0 >= 0 && s0 + nc <= src->k && d0 >= 0 && d0 + nc <= dst->k, "bad column windows");
    if (rows <= 0) rows = src->n;
    NEPB_CHECK_ARG(rows <= src->n && rows <= dst->n, "rows exceed the block height");
    const double2 a = alpha ? make_double2(alpha[0], alpha[1]) : make_double2(1.0, 0.0);
    NEPB_LAUNCH(copy_cols_kernel, (unsigned)((rows * nc + 255) / 256), 256, 0, rows, nc, (const double2*)src->d.p, src->k, s0, (double2*)dst->d.p,
                dst->k, d0, a);
    NEPB_LAUNCH_CHECK();
    return NEPB_OK;
}

}  // extern "C"
