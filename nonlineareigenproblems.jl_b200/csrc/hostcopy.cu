// Host <-> device movement of dense blocks for the host-buffer entry points of the C ABI.
//
// The reference's plugin calls take and return Julia arrays (compute_Mlincomb / compute_MM / lin_solve, src/NEPCore.jl:113-160,
// src/LinSolvers.jl:135-137): column-major n x k ComplexF64 in pageable memory.  On the device blocks are row-major.  A plain
// cudaMemcpy from pageable memory runs at 6-7 GB/s and serialises with the transposition; here the transfer is a pipeline over
// column chunks:
//     host threads copy chunk i+1 into a pinned slot  |  DMA of chunk i (pinned -> HBM, PCIe rate)  |  transposition of chunk i-1
// (and the mirror image on the way back).  Buffers the caller registered with nepb_host_register (or any pinned memory) skip
// the pinned hop and are DMA'd directly.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <omp.h>
#include <cstdlib>

#include "common.h"

namespace nepb {

// tiled transpositions between the host layout (column-major) and the device layout (row-major)
__global__ void __launch_bounds__(256) colmajor_to_rowmajor_kernel(int64_t n, int kc, const double2* __restrict__ src, int64_t lds,
                                                                   double2* __restrict__ dst, int ldd, int k0) {
    __shared__ double2 tile[32][33];
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = ty; j < 32; j += 8)
        if (r0 + tx < n && c0 + j < kc) tile[j][tx] = src[(size_t)(c0 + j) * lds + r0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (r0 + j < n && c0 + tx < kc) dst[(size_t)(r0 + j) * ldd + k0 + c0 + tx] = tile[tx][j];
}

__global__ void __launch_bounds__(256) rowmajor_to_colmajor_kernel(int64_t n, int kc, const double2* __restrict__ src, int lds, int k0,
                                                                   double2* __restrict__ dst, int64_t ldd) {
    __shared__ double2 tile[32][33];
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = ty; j < 32; j += 8)
        if (r0 + j < n && c0 + tx < kc) tile[j][tx] = src[(size_t)(r0 + j) * lds + k0 + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (r0 + tx < n && c0 + j < kc) dst[(size_t)(c0 + j) * ldd + r0 + tx] = tile[tx][j];
}


namespace {

constexpr int NSLOT = 3;
struct Stager {
    std::mutex mtx;
    int device = -1;
    char* pinned[NSLOT] = {nullptr, nullptr, nullptr};
    size_t slot_bytes = 0;
    cudaStream_t cs = nullptr;  // copy stream
    cudaEvent_t slot_ev[NSLOT] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    bool ready = false;
};
Stager g_st[16];

int stager_get(size_t want, Stager** out) {
    int dev = 0;
    NEPB_CUDA(cudaGetDevice(&dev));
    Stager& S = g_st[dev & 15];
    if (!S.ready) {
        S.device = dev;
        NEPB_CUDA(cudaStreamCreateWithFlags(&S.cs, cudaStreamNonBlocking));
        for (int i = 0; i < NSLOT; ++i) NEPB_CUDA(cudaEventCreateWithFlags(&S.slot_ev[i], cudaEventDisableTiming));
        NEPB_CUDA(cudaEventCreateWithFlags(&S.ev_in, cudaEventDisableTiming));
        NEPB_CUDA(cudaEventCreateWithFlags(&S.ev_out, cudaEventDisableTiming));
        S.ready = true;
    }
    if (want > S.slot_bytes) {
        NEPB_CUDA(cudaStreamSynchronize(S.cs));
        for (int i = 0; i < NSLOT; ++i) {
            if (S.pinned[i]) cudaFreeHost(S.pinned[i]);
            S.pinned[i] = nullptr;
        }
        S.slot_bytes = 0;
        for (int i = 0; i < NSLOT; ++i) NEPB_CUDA(cudaMallocHost((void**)&S.pinned[i], want));
        S.slot_bytes = want;
    }
    *out = &S;
    return NEPB_OK;
}

static const int g_copy_threads = [] {
    int t = 8;  // measured on the 16-core GPU box: 8 copy threads reach the PCIe rate, 16 oversubscribe (the caller and the driver need cores too)
    if (const char* e = getenv("NEPB_COPY_THREADS")) t = atoi(e);
    return std::max(1, std::min(t, (int)omp_get_num_procs()));
}();

// memcpy spread over the host threads (a single thread moves ~10 GB/s, a PCIe 5 x16 link ~50 GB/s)
void par_copy(char* dst, const char* src, size_t bytes) {
    constexpr size_t PIECE = (size_t)1 << 20;
    const int64_t np = (int64_t)((bytes + PIECE - 1) / PIECE);
    if (np <= 2) {
        memcpy(dst, src, bytes);
        return;
    }
#pragma omp parallel for schedule(static) num_threads(g_copy_threads)
    for (int64_t i = 0; i < np; ++i) {
        const size_t o = (size_t)i * PIECE;
        memcpy(dst + o, src + o, std::min(PIECE, bytes - o));
    }
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

constexpr size_t SLOT_TARGET = (size_t)16 << 20;

}  // namespace

// host column-major n x kc (leading dimension ld, complex elements) -> dst row-major, columns [k0, k0+kc) of rows of length ldd.
// `stage` is the caller's device staging buffer (column-major copy of the block).  The result is ordered on stream().
int upload_colmajor(int64_t n, int kc, const double* host, int64_t ld, DevBuf<double>& stage, double* dst, int ldd, int k0) {
    const size_t col_bytes = (size_t)n * 16;
    NEPB_CUDA(stage.reserve((size_t)2 * n * kc));
    if (col_bytes * kc <= (size_t)256 << 10) {  // small blocks: one copy (the driver stages it) + one transposition, on the caller's stream
        NEPB_CUDA(cudaMemcpy2DAsync(stage.p, col_bytes, host, (size_t)ld * 16, col_bytes, kc, cudaMemcpyHostToDevice, stream()));
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((kc + 31) / 32));
        NEPB_LAUNCH(colmajor_to_rowmajor_kernel, grid, 256, 0, n, kc, (const double2*)stage.p, n, (double2*)dst, ldd, k0);
        NEPB_LAUNCH_CHECK();
        return NEPB_OK;
    }
    const bool direct = is_pinned(host);
    const int cols_per_chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)kc, SLOT_TARGET / col_bytes));
    Stager* S = nullptr;
    int rc = stager_get(direct ? 1 : (size_t)cols_per_chunk * col_bytes, &S);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(S->mtx);
    // the copy stream must not overwrite `stage` / `dst` while earlier work of the caller's stream still uses them
    NEPB_CUDA(cudaEventRecord(S->ev_in, stream()));
    NEPB_CUDA(cudaStreamWaitEvent(S->cs, S->ev_in, 0));
    int i = 0;
    for (int c0 = 0; c0 < kc; c0 += cols_per_chunk, ++i) {
        const int cc = std::min(cols_per_chunk, kc - c0);
        double* sdst = stage.p + (size_t)2 * n * c0;
        if (direct) {
            NEPB_CUDA(cudaMemcpy2DAsync(sdst, col_bytes, host + (size_t)2 * ld * c0, (size_t)ld * 16, col_bytes, cc, cudaMemcpyHostToDevice, S->cs));
        } else {
            const int slot = i % NSLOT;
            NEPB_CUDA(cudaEventSynchronize(S->slot_ev[slot]));  // the DMA that last read this slot has finished
            for (int c = 0; c < cc; ++c)
                par_copy(S->pinned[slot] + (size_t)c * col_bytes, (const char*)(host + (size_t)2 * ld * (c0 + c)), col_bytes);
            NEPB_CUDA(cudaMemcpyAsync(sdst, S->pinned[slot], col_bytes * cc, cudaMemcpyHostToDevice, S->cs));
            NEPB_CUDA(cudaEventRecord(S->slot_ev[slot], S->cs));
        }
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((cc + 31) / 32));
        colmajor_to_rowmajor_kernel<<<grid, 256, 0, S->cs>>>(n, cc, (const double2*)sdst, n, (double2*)dst, ldd, k0 + c0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    NEPB_LAUNCH_CHECK();
    NEPB_CUDA(cudaEventRecord(S->ev_out, S->cs));
    NEPB_CUDA(cudaStreamWaitEvent(stream(), S->ev_out, 0));
    if (direct) NEPB_CUDA(cudaStreamSynchronize(S->cs));  // the caller may reuse its buffer as soon as we return
    return NEPB_OK;
}

// src row-major (rows of length lds), columns [k0, k0+kc) -> host column-major n x kc (leading dimension ld).  Returns when
// the host array is complete.
int download_colmajor(int64_t n, int kc, const double* src, int lds, int k0, DevBuf<double>& stage, double* host, int64_t ld) {
    const size_t col_bytes = (size_t)n * 16;
    NEPB_CUDA(stage.reserve((size_t)2 * n * kc));
    if (col_bytes * kc <= (size_t)256 << 10) {
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((kc + 31) / 32));
        NEPB_LAUNCH(rowmajor_to_colmajor_kernel, grid, 256, 0, n, kc, (const double2*)src, lds, k0, (double2*)stage.p, n);
        NEPB_LAUNCH_CHECK();
        NEPB_CUDA(cudaMemcpy2DAsync(host, (size_t)ld * 16, stage.p, col_bytes, col_bytes, kc, cudaMemcpyDeviceToHost, stream()));
        NEPB_CUDA(cudaStreamSynchronize(stream()));
        return NEPB_OK;
    }
    const bool direct = is_pinned(host);
    const int cols_per_chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)kc, SLOT_TARGET / col_bytes));
    Stager* S = nullptr;
    int rc = stager_get(direct ? 1 : (size_t)cols_per_chunk * col_bytes, &S);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(S->mtx);
    NEPB_CUDA(cudaEventRecord(S->ev_in, stream()));  // the producer of `src`
    NEPB_CUDA(cudaStreamWaitEvent(S->cs, S->ev_in, 0));
    int i = 0, pending = -1, pend_c0 = 0, pend_cc = 0;
    auto drain = [&](int slot, int c0, int cc) -> int {
        NEPB_CUDA(cudaEventSynchronize(S->slot_ev[slot]));
        for (int c = 0; c < cc; ++c)
            par_copy((char*)(host + (size_t)2 * ld * (c0 + c)), S->pinned[slot] + (size_t)c * col_bytes, col_bytes);
        return NEPB_OK;
    };
    for (int c0 = 0; c0 < kc; c0 += cols_per_chunk, ++i) {
        const int cc = std::min(cols_per_chunk, kc - c0);
        double* ssrc = stage.p + (size_t)2 * n * c0;
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((cc + 31) / 32));
        rowmajor_to_colmajor_kernel<<<grid, 256, 0, S->cs>>>(n, cc, (const double2*)src, lds, k0 + c0, (double2*)ssrc, n);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (direct) {
            NEPB_CUDA(cudaMemcpy2DAsync(host + (size_t)2 * ld * c0, (size_t)ld * 16, ssrc, col_bytes, col_bytes, cc, cudaMemcpyDeviceToHost, S->cs));
        } else {
            const int slot = i % NSLOT;
            // slot reuse is safe: the chunk that used it NSLOT iterations ago was drained below before we got here
            NEPB_CUDA(cudaMemcpyAsync(S->pinned[slot], ssrc, col_bytes * cc, cudaMemcpyDeviceToHost, S->cs));
            NEPB_CUDA(cudaEventRecord(S->slot_ev[slot], S->cs));
            if (pending >= 0) {
                rc = drain(pending, pend_c0, pend_cc);
                if (rc) return rc;
            }
            pending = slot;
            pend_c0 = c0;
            pend_cc = cc;
        }
    }
    NEPB_LAUNCH_CHECK();
    if (!direct && pending >= 0) {
        rc = drain(pending, pend_c0, pend_cc);
        if (rc) return rc;
    }
    NEPB_CUDA(cudaStreamSynchronize(S->cs));
    return NEPB_OK;
}

}  // namespace nepb

using namespace nepb;

extern "C" {

// Page-lock a host buffer the caller passes to the host-buffer entry points again and again (Krylov bases, probe / moment
// blocks): transfers from / to it then run as direct DMA at PCIe rate without the pinned hop.
int nepb_host_register(void* ptr, int64_t bytes) {
    NEPB_CHECK_ARG(ptr && bytes > 0, "bad arguments");
    NEPB_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return NEPB_OK;
}

int nepb_host_unregister(void* ptr) {
    NEPB_CHECK_ARG(ptr, "ptr is NULL");
    NEPB_CUDA(cudaHostUnregister(ptr));
    return NEPB_OK;
}

}  // extern "C"
