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

// ---------------------------------------------------------------------------------------------
// Round 2 ZGEMM: 128 rows x (up to) 56 columns per CTA, 8 warps, each warp 16 rows (2 DMMA row tiles) x ALL column tiles of the
// CTA (<= 7), so every A fragment feeds up to 7 products and the quantisation loss on the small dimension is q / (8 ceil(q/8))
// (50 -> 89 %, 100 -> 96 %, 200 -> 100 %; with 64-wide tiles it was 78 %, profiles/r2_ncu_gemm.txt: the DMMA pipe was 79 % busy).
// K in chunks of 16 through a 3-stage cp.async ring.  The MMA wants real and imaginary parts as separate FP64 operands: the
// 8-byte cp.async copies de-interleave on the way into shared memory (re plane / im plane), with row pitches of 20 (A) and 68
// (B) doubles so that the 16 lanes of a 64-bit shared-memory phase hit 16 different bank pairs.  Out-of-range rows / columns /
// k are zero-filled by the copy itself; column tiles that lie completely beyond q are skipped (CTA-uniform).
// ---------------------------------------------------------------------------------------------
constexpr int ZT_NT = 7, ZT_N = 8 * ZT_NT, ZT_K = 16, ZT_ST = 3;
constexpr int ZA_LD = ZT_K + 4, ZB_LD = 64 + 4;
__host__ __device__ constexpr int zt_stage_doubles(int zm) { return 2 * zm * ZA_LD + 2 * ZT_K * ZB_LD; }
__host__ __device__ constexpr size_t zt_smem(int zm) { return (size_t)ZT_ST * zt_stage_doubles(zm) * 8; }

__device__ __forceinline__ void cp_async_8z(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;  // src-size 0: the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gsrc), "r"(sz));
}

// ZT_M rows per CTA = 16 per warp: 128 (8 warps, one CTA per SM) or 64 (4 warps, two CTAs per SM: the ring waits / barriers of
// one CTA run under the MMAs of the other)
template <int ZT_M>
__global__ void __launch_bounds__(2 * ZT_M, 128 / ZT_M) block_gemm_dmma2_kernel(int64_t rows, int ka, int q, int cols_per_cta, const double2* __restrict__ A,
                                                                  int lda, const double2* __restrict__ C /* row-major ka x q */,
                                                                  double2* __restrict__ Y, int ldy) {
    extern __shared__ double zsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * ZT_M;
    const int col0 = blockIdx.y * cols_per_cta;
    const int ncols = min(cols_per_cta, q - col0);  // columns of this CTA (<= 56)
    const int nt_act = (ncols + 7) / 8;
    const int nk = (ka + ZT_K - 1) / ZT_K;
    constexpr int ZT_STAGE_DOUBLES = zt_stage_doubles(ZT_M), NTH = 2 * ZT_M;
    auto stage_ptr = [&](int st) { return zsm + (size_t)st * ZT_STAGE_DOUBLES; };
    auto load_chunk = [&](int kc, int st) {
        double* sAr = stage_ptr(st);
        double* sAi = sAr + ZT_M * ZA_LD;
        double* sBr = sAi + ZT_M * ZA_LD;
        double* sBi = sBr + ZT_K * ZB_LD;
        const int k0 = kc * ZT_K;
        for (int idx = tid; idx < ZT_M * ZT_K; idx += NTH) {  // A tile: 128 rows x 16 k, 256 contiguous bytes per row
            const int r = idx / ZT_K, kk = idx % ZT_K;
            const bool ok = row0 + r < rows && k0 + kk < ka;
            const double* src = (const double*)(A + (ok ? (size_t)(row0 + r) * lda + k0 + kk : 0));
            cp_async_8z(sAr + r * ZA_LD + kk, src, ok);
            cp_async_8z(sAi + r * ZA_LD + kk, src + 1, ok);
        }
        for (int idx = tid; idx < ZT_K * ZT_N; idx += NTH) {  // B tile: 16 k x 56 columns of the small matrix
            const int kk = idx / ZT_N, c = idx % ZT_N;
            const bool ok = k0 + kk < ka && c < ncols;
            const double* src = (const double*)(C + (ok ? (size_t)(k0 + kk) * q + col0 + c : 0));
            cp_async_8z(sBr + kk * ZB_LD + c, src, ok);
            cp_async_8z(sBi + kk * ZB_LD + c, src + 1, ok);
        }
    };
    double cr[2][ZT_NT][2], ci[2][ZT_NT][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < ZT_NT; ++nt) cr[mt][nt][0] = cr[mt][nt][1] = ci[mt][nt][0] = ci[mt][nt][1] = 0.0;
    for (int st = 0; st < ZT_ST - 1; ++st) {
        if (st < nk) load_chunk(st, st);
        asm volatile("cp.async.commit_group;" ::);
    }
    const int ar = lane >> 2, ak = lane & 3;
    for (int kc = 0; kc < nk; ++kc) {
        asm volatile("cp.async.wait_group %0;" ::"n"(ZT_ST - 2));
        __syncthreads();  // chunk kc has landed for every thread, and everybody is done with the stage refilled below
        if (kc + ZT_ST - 1 < nk) load_chunk(kc + ZT_ST - 1, (kc + ZT_ST - 1) % ZT_ST);
        asm volatile("cp.async.commit_group;" ::);
        const double* sAr = stage_ptr(kc % ZT_ST);
        const double* sAi = sAr + ZT_M * ZA_LD;
        const double* sBr = sAi + ZT_M * ZA_LD;
        const double* sBi = sBr + ZT_K * ZB_LD;
#pragma unroll
        for (int k4 = 0; k4 < ZT_K; k4 += 4) {
            double a_r[2], a_i[2], na_i[2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int r = warp * 16 + mt * 8 + ar;
                a_r[mt] = sAr[r * ZA_LD + k4 + ak];
                a_i[mt] = sAi[r * ZA_LD + k4 + ak];
                na_i[mt] = -a_i[mt];
            }
#pragma unroll
            for (int nt = 0; nt < ZT_NT; ++nt) {
                if (nt < nt_act) {  // CTA-uniform
                    const double b_r = sBr[(k4 + ak) * ZB_LD + nt * 8 + ar];
                    const double b_i = sBi[(k4 + ak) * ZB_LD + nt * 8 + ar];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        dmma(cr[mt][nt][0], cr[mt][nt][1], a_r[mt], b_r);
                        dmma(ci[mt][nt][0], ci[mt][nt][1], a_r[mt], b_i);
                        dmma(cr[mt][nt][0], cr[mt][nt][1], na_i[mt], b_i);
                        dmma(ci[mt][nt][0], ci[mt][nt][1], a_i[mt], b_r);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int64_t r = row0 + warp * 16 + mt * 8 + ar;
        if (r >= rows) continue;
#pragma unroll
        for (int nt = 0; nt < ZT_NT; ++nt) {
            const int c = nt * 8 + 2 * ak;
            if (c < ncols) Y[(size_t)r * ldy + col0 + c] = make_double2(cr[mt][nt][0], ci[mt][nt][0]);
            if (c + 1 < ncols) Y[(size_t)r * ldy + col0 + c + 1] = make_double2(cr[mt][nt][1], ci[mt][nt][1]);
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

// ---------------------------------------------------------------------------------------------
// Round 2: one Gram-Schmidt sweep = TWO launches and no host round trip.  Each warp streams four rows at a time (all loads
// issued before the arithmetic: at k = 50 a row is only 800 bytes and one row per warp in flight left the kernel latency
// bound, profiles/r1_c5_tiar_dense_blocks.txt); the CTA partial sums are reduced in CTA order by the LAST CTA to finish
// (ticket counter + __threadfence), so the result stays bitwise reproducible; the DGKS decision ||w|| < ||h||/sqrt(2)
// (IterativeSolvers orthogonalize.jl) is taken on the device: later sweeps are enqueued up front and return at once when the
// flag says they are not needed.
// ctl[0] = ||h||^2 of the last sweep, ctl[1] = ||w||^2, ctl[2] = sweeps done, ctl[3] = continue flag
// ---------------------------------------------------------------------------------------------
constexpr int ORTH_UR = 4;

// last CTA to finish (ticket) reduces the CTA partial sums in CTA order: h of this sweep, accumulated h, ||h||^2
__device__ __forceinline__ void orth_dot_finish(int k, double2* __restrict__ partial, const double2* myp /* shared, k entries */,
                                                double2* __restrict__ hcur, double2* __restrict__ hsum, int accumulate, double* __restrict__ ctl,
                                                unsigned* __restrict__ ticket, double* sn, unsigned* last) {
    for (int j = threadIdx.x; j < k; j += 256) partial[(size_t)blockIdx.x * k + j] = myp[j];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!*last) return;
    __threadfence();
    double nn = 0.0;
    for (int j = threadIdx.x; j < k; j += 256) {
        double2 t = make_double2(0.0, 0.0);
#pragma unroll 8
        for (unsigned c = 0; c < gridDim.x; ++c) {
            const double2 pv = __ldcg(partial + (size_t)c * k + j);
            t.x += pv.x;
            t.y += pv.y;
        }
        hcur[j] = t;
        if (accumulate) {
            hsum[j].x += t.x;
            hsum[j].y += t.y;
        } else {
            hsum[j] = t;
        }
        nn += t.x * t.x + t.y * t.y;
    }
    sn[threadIdx.x] = nn;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 256; ++i) t += sn[i];
        ctl[0] = t;
        *ticket = 0u;
    }
}

// thread 0 of the last CTA: ||w||^2, sweep count and the DGKS decision
__device__ __forceinline__ void orth_update_finish(int k, double cta_sum, double* __restrict__ npartial, double* __restrict__ ctl,
                                                   unsigned* __restrict__ ticket) {
    npartial[blockIdx.x] = cta_sum;
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        __threadfence();
        double tot = 0.0;
        for (unsigned c = 0; c < gridDim.x; ++c) tot += __ldcg(npartial + c);
        ctl[1] = tot;
        ctl[2] += 1.0;
        // DGKS: another sweep while ||w|| < ||h|| / sqrt(2) (and there is something to project against)
        ctl[3] = (k > 0 && sqrt(tot) < 0.7071067811865476 * sqrt(ctl[0])) ? 1.0 : 0.0;
        *ticket = 0u;
    }
}

// wide bases (k > 128): lane = column, a warp per row
template <int CPL>
__global__ void __launch_bounds__(256) orth_dot_fused_kernel(int64_t rows, int k, const double2* __restrict__ V, int ldv, const double2* __restrict__ w,
                                                             int ldw, double2* __restrict__ partial, double2* __restrict__ hcur,
                                                             double2* __restrict__ hsum, int accumulate, double* __restrict__ ctl,
                                                             unsigned* __restrict__ ticket, int need_flag) {
    if (need_flag && ctl[3] == 0.0) return;
    __shared__ double2 red[8][32 * CPL];
    __shared__ double sn[256];
    __shared__ unsigned last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = make_double2(0.0, 0.0);
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    constexpr int UR = CPL >= 8 ? 1 : 4;  // rows in flight per warp (a 100-column row is only 1.6 KB)
    for (int64_t rb = r0 + warp; rb < r1; rb += 8 * UR) {
        double2 wr[UR], v[UR][CPL];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
            const int64_t r = rb + 8 * u;
            wr[u] = (r < r1) ? w[(size_t)r * ldw] : make_double2(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int j = lane + 32 * c;
                v[u][c] = (r < r1 && j < k) ? V[(size_t)r * ldv + j] : make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < UR; ++u)
#pragma unroll
            for (int c = 0; c < CPL; ++c) {  // conj(v) * w
                acc[c].x = fma(v[u][c].x, wr[u].x, acc[c].x);
                acc[c].x = fma(v[u][c].y, wr[u].y, acc[c].x);
                acc[c].y = fma(v[u][c].x, wr[u].y, acc[c].y);
                acc[c].y = fma(-v[u][c].y, wr[u].x, acc[c].y);
            }
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) red[warp][lane + 32 * c] = acc[c];
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += 256) {
        double2 t = red[0][j];
        for (int q = 1; q < 8; ++q) {
            t.x += red[q][j].x;
            t.y += red[q][j].y;
        }
        red[0][j] = t;
    }
    __syncthreads();
    orth_dot_finish(k, partial, red[0], hcur, hsum, accumulate, ctl, ticket, sn, &last);
}

template <int CPL>
__global__ void __launch_bounds__(256) orth_update_fused_kernel(int64_t rows, int k, const double2* __restrict__ V, int ldv, double2* __restrict__ w,
                                                                int ldw, const double2* __restrict__ h, double* __restrict__ npartial,
                                                                double* __restrict__ ctl, unsigned* __restrict__ ticket, int need_flag) {
    if (need_flag && ctl[3] == 0.0) return;
    __shared__ double2 sh[32 * CPL];
    __shared__ double sn[8];
    for (int j = threadIdx.x; j < 32 * CPL; j += 256) sh[j] = j < k ? h[j] : make_double2(0.0, 0.0);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    double nn = 0.0;
    constexpr int UR = CPL >= 8 ? 1 : 4;
    for (int64_t rb = r0 + warp; rb < r1; rb += 8 * UR) {  // warp-uniform
        double2 v[UR][CPL];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
            const int64_t r = rb + 8 * u;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int j = lane + 32 * c;
                v[u][c] = (r < r1 && j < k) ? V[(size_t)r * ldv + j] : make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < UR; ++u) {
            double2 t = make_double2(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const double2 hj = sh[lane + 32 * c];
                t.x = fma(v[u][c].x, hj.x, t.x);
                t.x = fma(-v[u][c].y, hj.y, t.x);
                t.y = fma(v[u][c].x, hj.y, t.y);
                t.y = fma(v[u][c].y, hj.x, t.y);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                t.x += __shfl_xor_sync(0xffffffffu, t.x, off);
                t.y += __shfl_xor_sync(0xffffffffu, t.y, off);
            }
            const int64_t r = rb + 8 * u;
            if (lane == 0 && r < r1) {
                double2 wr = w[(size_t)r * ldw];
                wr.x -= t.x;
                wr.y -= t.y;
                w[(size_t)r * ldw] = wr;
                nn += wr.x * wr.x + wr.y * wr.y;
            }
        }
    }
    if (lane == 0) sn[warp] = nn;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += sn[q];
        orth_update_finish(k, t, npartial, ctl, ticket);
    }
}

// narrow bases (k <= 128): 8 lanes per row, lane g owns the columns g, g + 8, ...  A warp instruction reads four rows with
// 128 contiguous bytes each, the row-wise reduction of the update needs 3 shuffle steps instead of 5, and 4 x UR rows are in
// flight per warp.
template <int CP8, int UR>
__global__ void __launch_bounds__(256) orth_dot_g8_kernel(int64_t rows, int k, const double2* __restrict__ V, int ldv, const double2* __restrict__ w,
                                                          int ldw, double2* __restrict__ partial, double2* __restrict__ hcur,
                                                          double2* __restrict__ hsum, int accumulate, double* __restrict__ ctl,
                                                          unsigned* __restrict__ ticket, int need_flag) {
    if (need_flag && ctl[3] == 0.0) return;
    __shared__ double2 red[8][8 * CP8];
    __shared__ double sn[256];
    __shared__ unsigned last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane & 7, rg = lane >> 3;
    double2 acc[CP8];
#pragma unroll
    for (int c = 0; c < CP8; ++c) acc[c] = make_double2(0.0, 0.0);
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    for (int64_t base = r0 + warp * 4; base < r1; base += 32 * UR) {  // warp-uniform trip count
        const int64_t rb = base + rg;
        double2 wr[UR], v[UR][CP8];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
            const int64_t r = rb + 32 * u;
            wr[u] = (r < r1) ? w[(size_t)r * ldw] : make_double2(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < CP8; ++c) {
                const int j = g + 8 * c;
                v[u][c] = (r < r1 && j < k) ? V[(size_t)r * ldv + j] : make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < UR; ++u)
#pragma unroll
            for (int c = 0; c < CP8; ++c) {  // conj(v) * w
                acc[c].x = fma(v[u][c].x, wr[u].x, acc[c].x);
                acc[c].x = fma(v[u][c].y, wr[u].y, acc[c].x);
                acc[c].y = fma(v[u][c].x, wr[u].y, acc[c].y);
                acc[c].y = fma(-v[u][c].y, wr[u].x, acc[c].y);
            }
    }
#pragma unroll
    for (int c = 0; c < CP8; ++c) {  // the four row groups of the warp
        acc[c].x += __shfl_xor_sync(0xffffffffu, acc[c].x, 8);
        acc[c].y += __shfl_xor_sync(0xffffffffu, acc[c].y, 8);
        acc[c].x += __shfl_xor_sync(0xffffffffu, acc[c].x, 16);
        acc[c].y += __shfl_xor_sync(0xffffffffu, acc[c].y, 16);
    }
    if (rg == 0) {
#pragma unroll
        for (int c = 0; c < CP8; ++c) red[warp][g + 8 * c] = acc[c];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += 256) {
        double2 t = red[0][j];
        for (int q = 1; q < 8; ++q) {
            t.x += red[q][j].x;
            t.y += red[q][j].y;
        }
        red[0][j] = t;
    }
    __syncthreads();
    orth_dot_finish(k, partial, red[0], hcur, hsum, accumulate, ctl, ticket, sn, &last);
}

template <int CP8, int UR>
__global__ void __launch_bounds__(256) orth_update_g8_kernel(int64_t rows, int k, const double2* __restrict__ V, int ldv, double2* __restrict__ w,
                                                             int ldw, const double2* __restrict__ h, double* __restrict__ npartial,
                                                             double* __restrict__ ctl, unsigned* __restrict__ ticket, int need_flag) {
    if (need_flag && ctl[3] == 0.0) return;
    __shared__ double sn[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane & 7, rg = lane >> 3;
    double2 hh[CP8];  // this lane's coefficients stay in registers
#pragma unroll
    for (int c = 0; c < CP8; ++c) hh[c] = (g + 8 * c < k) ? h[g + 8 * c] : make_double2(0.0, 0.0);
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    double nn = 0.0;
    for (int64_t base = r0 + warp * 4; base < r1; base += 32 * UR) {  // warp-uniform trip count: shuffles inside
        const int64_t rb = base + rg;
        double2 v[UR][CP8], wr[UR];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
            const int64_t r = rb + 32 * u;
            wr[u] = (r < r1 && g == 0) ? w[(size_t)r * ldw] : make_double2(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < CP8; ++c) {
                const int j = g + 8 * c;
                v[u][c] = (r < r1 && j < k) ? V[(size_t)r * ldv + j] : make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < UR; ++u) {
            double2 t = make_double2(0.0, 0.0);
#pragma unroll
            for (int c = 0; c < CP8; ++c) {
                t.x = fma(v[u][c].x, hh[c].x, t.x);
                t.x = fma(-v[u][c].y, hh[c].y, t.x);
                t.y = fma(v[u][c].x, hh[c].y, t.y);
                t.y = fma(v[u][c].y, hh[c].x, t.y);
            }
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {
                t.x += __shfl_xor_sync(0xffffffffu, t.x, off);
                t.y += __shfl_xor_sync(0xffffffffu, t.y, off);
            }
            const int64_t r = rb + 32 * u;
            if (g == 0 && r < r1) {
                wr[u].x -= t.x;
                wr[u].y -= t.y;
                w[(size_t)r * ldw] = wr[u];
                nn += wr[u].x * wr[u].x + wr[u].y * wr[u].y;
            }
        }
    }
    // fixed-order sum over the row groups of the warp, then over the warps
    nn += __shfl_xor_sync(0xffffffffu, nn, 8);
    nn += __shfl_xor_sync(0xffffffffu, nn, 16);
    if (lane == 0) sn[warp] = nn;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += sn[q];
        orth_update_finish(k, t, npartial, ctl, ticket);
    }
}

__global__ void __launch_bounds__(256) scale_col_dev_kernel(int64_t rows, double2* __restrict__ w, int ldw, const double* __restrict__ ctl) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double n2 = ctl[1];
    const double alpha = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
    double2 v = w[(size_t)r * ldw];
    v.x *= alpha;
    v.y *= alpha;
    w[(size_t)r * ldw] = v;
}

struct OrthScratch {
    DevBuf<double2> partial, hcur, hsum;
    DevBuf<double> npartial, scal;
    DevBuf<unsigned> ticket;
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
    const int ncta = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)sm_count() * (k <= 64 ? 2 : 4));  // measured per mapping (profiles/r2_c5_tiar_dense_blocks.txt)
    const int kk = std::max(k, 1);
    NEPB_CUDA(g_orth.partial.reserve((size_t)ncta * kk));
    NEPB_CUDA(g_orth.hcur.reserve(kk));
    NEPB_CUDA(g_orth.hsum.reserve(kk));
    NEPB_CUDA(g_orth.npartial.reserve(ncta));
    NEPB_CUDA(g_orth.scal.reserve(4));
    if (!g_orth.ticket.p) {
        NEPB_CUDA(g_orth.ticket.alloc(2));
        NEPB_CUDA(cudaMemsetAsync(g_orth.ticket.p, 0, 2 * sizeof(unsigned), stream()));
    }
    const double2* Vp = (const double2*)V->d.p;
    double2* wp = (double2*)W->d.p + wcol;
    const int ldv = V->k, ldw = W->k;
    double ctl[4] = {0.0, 0.0, 0.0, 0.0};  // {||h||^2 of the last sweep, ||w||^2, sweeps, continue}
    NEPB_CUDA(cudaMemsetAsync(g_orth.scal.p, 0, 4 * sizeof(double), stream()));
    if (k > 128) {
        // wide bases keep the round-1 sequence (dot, fold, update, fold + one read-back per sweep): at k = 200 the sweep is
        // bandwidth bound either way and this form measured faster (1.27 vs 1.71 ms, profiles/r2_c5_tiar_dense_blocks.txt)
        double hs[2] = {0.0, 0.0};
        int nsweep = 0;
        for (;;) {
            NEPB_LAUNCH(orth_dot_kernel, ncta, 256, 0, rows, k, Vp, ldv, (const double2*)wp, ldw, g_orth.partial.p);
            NEPB_LAUNCH(orth_reduce_kernel, 1, 256, 0, ncta, k, (const double2*)g_orth.partial.p, g_orth.hcur.p, g_orth.hsum.p, nsweep > 0 ? 1 : 0,
                        g_orth.scal.p);
            NEPB_LAUNCH(orth_update_kernel, ncta, 256, 0, rows, k, Vp, ldv, wp, ldw, (const double2*)g_orth.hcur.p, g_orth.npartial.p);
            NEPB_LAUNCH(orth_norm_reduce_kernel, 1, 32, 0, ncta, (const double*)g_orth.npartial.p, g_orth.scal.p + 1);
            NEPB_LAUNCH_CHECK();
            NEPB_CUDA(cudaMemcpyAsync(hs, g_orth.scal.p, sizeof(hs), cudaMemcpyDeviceToHost, stream()));
            NEPB_CUDA(cudaStreamSynchronize(stream()));
            ++nsweep;
            if (!(std::sqrt(hs[1]) < 0.7071067811865476 * std::sqrt(hs[0])) || nsweep >= 8) break;
        }
        ctl[1] = hs[1];
        ctl[2] = nsweep;
        NEPB_CUDA(cudaMemcpyAsync(h, g_orth.hsum.p, sizeof(double) * 2 * k, cudaMemcpyDeviceToHost, stream()));
        NEPB_LAUNCH(scale_col_dev_kernel, (unsigned)((rows + 255) / 256), 256, 0, rows, wp, ldw, (const double*)g_orth.scal.p);
        NEPB_LAUNCH_CHECK();
        NEPB_CUDA(cudaStreamSynchronize(stream()));
        *nrm_out = std::sqrt(ctl[1]);
        if (sweeps) *sweeps = nsweep;
        return NEPB_OK;
    }
    // up to MAXSW sweeps enqueued back to back; sweeps after the first return immediately unless the device-side DGKS test
    // (orth_update_fused_kernel) asked for them -- no host round trip inside the orthogonalisation
    constexpr int MAXSW = 4;
    for (int sw = 0; sw < MAXSW; ++sw) {
#define NEPB_ORTH_SWEEP(DOT_, UPD_)                                                                                                         \
    do {                                                                                                                                   \
        if (k > 0)                                                                                                                         \
            NEPB_LAUNCH(DOT_, ncta, 256, 0, rows, k, Vp, ldv, (const double2*)wp, ldw, g_orth.partial.p, g_orth.hcur.p, g_orth.hsum.p,      \
                        sw > 0 ? 1 : 0, g_orth.scal.p, g_orth.ticket.p, sw > 0 ? 1 : 0);                                                    \
        NEPB_LAUNCH(UPD_, ncta, 256, 0, rows, k, Vp, ldv, wp, ldw, (const double2*)g_orth.hcur.p, g_orth.npartial.p, g_orth.scal.p,          \
                    g_orth.ticket.p + 1, sw > 0 ? 1 : 0);                                                                                   \
    } while (0)
        if (k <= 32) NEPB_ORTH_SWEEP((orth_dot_g8_kernel<4, 4>), (orth_update_g8_kernel<4, 4>));
        else if (k <= 64) NEPB_ORTH_SWEEP((orth_dot_g8_kernel<8, 2>), (orth_update_g8_kernel<8, 2>));
        else NEPB_ORTH_SWEEP((orth_dot_fused_kernel<4>), (orth_update_fused_kernel<4>));  // k <= 128; measured: beats <16, 1> of the g8 form
#undef NEPB_ORTH_SWEEP
        if (k == 0) break;
    }
    NEPB_LAUNCH(scale_col_dev_kernel, (unsigned)((rows + 255) / 256), 256, 0, rows, wp, ldw, (const double*)g_orth.scal.p);
    NEPB_LAUNCH_CHECK();
    NEPB_CUDA(cudaMemcpyAsync(ctl, g_orth.scal.p, sizeof(ctl), cudaMemcpyDeviceToHost, stream()));
    if (k > 0) NEPB_CUDA(cudaMemcpyAsync(h, g_orth.hsum.p, sizeof(double) * 2 * k, cudaMemcpyDeviceToHost, stream()));
    NEPB_CUDA(cudaStreamSynchronize(stream()));
    *nrm_out = std::sqrt(ctl[1]);
    if (sweeps) *sweeps = (int)ctl[2];
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
    static const bool old_kernel = getenv("NEPB_GEMM_V1") && atoi(getenv("NEPB_GEMM_V1")) != 0;
    if (old_kernel) {
        dim3 grid((unsigned)((rows + GM_R - 1) / GM_R), (unsigned)((q + GM_C - 1) / GM_C));
        NEPB_LAUNCH(block_gemm_dmma_kernel, grid, 128, 0, rows, ka, q, (const double2*)A->d.p + acol0, A->k, (const double2*)g_gemm_c.p,
                    (double2*)Y->d.p + ycol0, Y->k);
    } else {
        static bool attr[16] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attr[dev & 15]) {
            NEPB_CUDA(cudaFuncSetAttribute(block_gemm_dmma2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zt_smem(128)));
            NEPB_CUDA(cudaFuncSetAttribute(block_gemm_dmma2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zt_smem(64)));
            attr[dev & 15] = true;
        }
        static const int zm = (getenv("NEPB_GEMM_ZM") && atoi(getenv("NEPB_GEMM_ZM")) == 64) ? 64 : 128;  // 64-row CTAs (two per SM) measured slower: 22.2 vs 25.6 TFLOP/s at k = 200
        // column tiles of equal width (a multiple of 8, at most 56): 50 -> 56, 100 -> 56 + 48, 200 -> 4 x 48 + 8 ... the tiles that
        // remain partly empty skip their empty 8-column blocks
        const int ntiles8 = (q + 7) / 8;
        const int nct = (ntiles8 + ZT_NT - 1) / ZT_NT;
        const int cols_per_cta = 8 * ((ntiles8 + nct - 1) / nct);
        dim3 grid((unsigned)((rows + zm - 1) / zm), (unsigned)nct);
        if (zm == 128)
            NEPB_LAUNCH(block_gemm_dmma2_kernel<128>, grid, 256, zt_smem(128), rows, ka, q, cols_per_cta, (const double2*)A->d.p + acol0, A->k,
                        (const double2*)g_gemm_c.p, (double2*)Y->d.p + ycol0, Y->k);
        else
            NEPB_LAUNCH(block_gemm_dmma2_kernel<64>, grid, 128, zt_smem(64), rows, ka, q, cols_per_cta, (const double2*)A->d.p + acol0, A->k,
                        (const double2*)g_gemm_c.p, (double2*)Y->d.p + ycol0, Y->k);
    }
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
