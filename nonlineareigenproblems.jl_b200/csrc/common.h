// Shared runtime bits of libnepb200: error reporting, the library stream, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <atomic>
#include <vector>

#include "../../include/nepb200.h"
#include "lu_symbolic.h"

namespace nepb {

void set_error(const char* fmt, ...);
cudaStream_t stream();       // stream the calling thread currently enqueues on
cudaStream_t main_stream();  // the library stream (timer events, host-facing copies)
void set_current_stream(cudaStream_t s);
void reset_current_stream();
extern std::atomic<int64_t> g_launches;
int sm_count();

#define NEPB_CHECK_ARG(cond, ...)                \
    do {                                         \
        if (!(cond)) {                           \
            nepb::set_error(__VA_ARGS__);        \
            return NEPB_E_INVALID;               \
        }                                        \
    } while (0)

#define NEPB_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            nepb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                            cudaGetErrorString(e__));                                           \
            return (e__ == cudaErrorMemoryAllocation) ? NEPB_E_NOMEM : NEPB_E_CUDA;             \
        }                                                                                       \
    } while (0)

// every kernel launch goes through this so that nepb_launch_count() is exact
#define NEPB_LAUNCH(kernel, grid, block, smem, ...)                        \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), nepb::stream()>>>(__VA_ARGS__);  \
        nepb::g_launches.fetch_add(1, std::memory_order_relaxed);          \
    } while (0)

#define NEPB_LAUNCH_CHECK() NEPB_CUDA(cudaGetLastError())

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    // grow-only scratch
    cudaError_t reserve(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
};

struct cplx {
    double re, im;
};

}  // namespace nepb

// ---- handle layouts (internal) ----------------------------------------------------------------
struct nepb_block {
    int64_t n = 0;
    int k = 0;
    nepb::DevBuf<double> d;  // row-major n x k complex interleaved: d[(r*k + c)*2 + {0,1}]
};

struct nepb_spmf {
    int64_t n = 0;
    int p = 0;
    int64_t nnz = 0;  // union pattern
    int is_complex = 0;
    int index_base = 1;
    int vw = 0;  // doubles per nonzero = p * (is_complex ? 2 : 1)
    // host copies of the integer structure (setup-time only, also used by the LU analysis)
    int64_t* h_colptr = nullptr;   // CSC union, 0-based, n+1
    int32_t* h_rowval = nullptr;   // CSC union, 0-based
    int32_t* h_rowptr = nullptr;   // CSR union
    int32_t* h_colind = nullptr;   // CSR union
    int32_t* h_csr_of_csc = nullptr;
    // device
    nepb::DevBuf<int32_t> d_rowptr, d_colind, d_csr_of_csc;
    nepb::DevBuf<double> d_vals;  // [nnz][vw], CSR order
    // scratch for host-facing calls
    mutable nepb::DevBuf<double> d_tmp_in, d_tmp_out, d_tmp_x, d_coef;
    mutable nepb::DevBuf<double> d_stage;
    // row tiles of the multi-column product (lazy, spmf.cu): <= R consecutive rows (R = 32 or 16) whose distinct column
    // indices (<= 6R, sorted) are staged once per tile in shared memory; lidx[e] = position of nonzero e's column in its
    // tile's list.  tile = two int4: (first row, rows, first entry of cols, distinct columns), (first nonzero, nonzeros, -, -)
    struct TileSet {
        int state = 0;  // 0 = not built, 1 = ready, -1 = not applicable (a row exceeds the tile budget) / no memory
        int64_t ntiles = 0, cols_total = 0, runs_total = 0;
        int max_cols = 0, max_nnz = 0;
        nepb::DevBuf<int4> tiles;
        nepb::DevBuf<int32_t> cols;
        nepb::DevBuf<uint16_t> lidx;
        nepb::DevBuf<int2> runs;  // runs of consecutive columns of every tile: (first column, position in the tile | length << 16)
    };
    mutable TileSet tiling[2];  // [0]: 32-row tiles, [1]: 16-row tiles
    // two-dimensional tiles (spmf.cu, round 2): a tile is S segments of <= R consecutive rows, one segment per "grid line" (rows
    // a dominant column offset apart), so that the tile's rows share most of their columns and far fewer rows of V are staged
    // per matrix row.  desc: (2 + S) int4 per tile: (first entry of cols, distinct columns, first run, runs),
    // (first entry of lidx, nonzeros, segments, -), then per segment (first row, rows, first nonzero, nonzeros).
    // lidx is stored tile by tile (padded to 8 entries) so that a tile's indices are one aligned bulk copy.
    struct TileSet2D {
        int state = 0;  // 0 = not built, 1 = ready, -1 = not applicable
        int S = 0, R = 0, line = 0;
        int64_t ntiles = 0;
        int max_cols = 0, max_nnz = 0;
        nepb::DevBuf<int4> desc;
        nepb::DevBuf<int32_t> cols;
        nepb::DevBuf<uint16_t> lidx;
        nepb::DevBuf<int2> runs;
    };
    mutable TileSet2D tiling2d;
    void* lu_symbolic = nullptr;  // owned by lu.cu (lazy)
    std::vector<void*> lu_matched;   // analyses of row-permuted patterns (static pivoting), newest last; owned by lu.cu
    bool lu_prefer_matched = false;  // a factorisation on the plain pattern met zero / tiny pivots
    nepb::LuOptions lu_opt;
    bool lu_opt_set = false;
    std::vector<int32_t> lu_user_perm;
    ~nepb_spmf();
};
