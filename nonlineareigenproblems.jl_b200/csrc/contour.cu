// Contour-integral quadrature on the device, sharded over ranks, with one NCCL reduce of the moment block.
//
// Replaces the N-point loop of integrate_interval(MatrixTrapezoidal, ...) (reference src/method_contour_common.jl:61-94)
// with the integrand of contour_beyn / contour_block_SS (src/method_beyncontour.jl:89-98, src/method_block_SS.jl:81-86):
//     S[:,:,j] = sum_i w[i,j] * M(lambda_i)^-1 Vh .
// Every rank owns a subset of the quadrature nodes; nodes are processed in batches: one batched multifrontal
// factorisation, one batched multi-RHS solve with the shared probe block Vh, one accumulate kernel.  The per-rank
// partial moments stay in HBM and are summed in place with a single ncclAllReduce (the reference's docs-only
// `@distributed (+)`, docs/src/tutorial_contour.md:205-218); NCCL is loaded at run time so that single-GPU use
// does not depend on it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.h"
#include "lu_internal.h"

namespace nepb {

__global__ void contour_accumulate_kernel(size_t nk, int nb, int mg, const double2* __restrict__ X, size_t x_stride,
                                          const double2* __restrict__ wgt, double2* __restrict__ S);
int upload_colmajor(int64_t n, int kc, const double* host, int64_t ld, DevBuf<double>& stage, double* dst, int ldd, int k0);
int download_colmajor(int64_t n, int kc, const double* src, int lds, int k0, DevBuf<double>& stage, double* host, int64_t ld);
int lu_fetch_info(nepb_lu* lu);

// ---- NCCL through dlopen ------------------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
    void* so = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_nranks = 1;

static int nccl_load() {
    if (g_nccl.so) return NEPB_OK;
    const char* cands[] = {getenv("NEPB_NCCL_LIB"), "libnccl.so.2",
                           "/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib/libnccl.so.2",
                           "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so"};
    void* so = nullptr;
    for (const char* c : cands) {
        if (!c) continue;
        so = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (so) break;
    }
    if (!so) {
        set_error("cannot load libnccl.so.2 (set NEPB_NCCL_LIB): %s", dlerror());
        return NEPB_E_UNSUPPORTED;
    }
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(so, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(so, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(so, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(so, "ncclAllReduce");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(so, "ncclGetErrorString");
    g_nccl.GetVersion = (int (*)(int*))dlsym(so, "ncclGetVersion");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.GetErrorString) {
        set_error("libnccl is missing a required symbol");
        return NEPB_E_UNSUPPORTED;
    }
    g_nccl.so = so;
    return NEPB_OK;
}

#define NEPB_NCCL(call)                                                                              \
    do {                                                                                             \
        int r__ = (call);                                                                            \
        if (r__ != 0) {                                                                              \
            set_error("NCCL error %d at %s:%d: %s", r__, __FILE__, __LINE__, g_nccl.GetErrorString(r__)); \
            return NEPB_E_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

}  // namespace nepb

using namespace nepb;

extern "C" {

int nepb_comm_unique_id(char id[128]) {
    NEPB_CHECK_ARG(id, "id is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId u;
    NEPB_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return NEPB_OK;
}

int nepb_comm_init(int nranks, int rank, const char id[128]) {
    NEPB_CHECK_ARG(id && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
    int rc = nccl_load();
    if (rc) return rc;
    if (g_comm) {
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    NEPB_NCCL(g_nccl.CommInitRank(&g_comm, nranks, u, rank));
    g_rank = rank;
    g_nranks = nranks;
    return NEPB_OK;
}

int nepb_comm_destroy(void) {
    if (g_comm) {
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_rank = 0;
    g_nranks = 1;
    return NEPB_OK;
}

int nepb_comm_info(int* nranks, int* rank, int* nccl_version) {
    if (nranks) *nranks = g_nranks;
    if (rank) *rank = g_rank;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.so && g_nccl.GetVersion) g_nccl.GetVersion(nccl_version);
    }
    return NEPB_OK;
}

// in-place sum over all ranks of a device buffer of `count` doubles (the moment block); no-op without a communicator
int nepb_comm_allreduce_sum_dev(void* dev_ptr, int64_t count) {
    NEPB_CHECK_ARG(dev_ptr && count >= 0, "bad arguments");
    if (!g_comm || g_nranks == 1) return NEPB_OK;
    NEPB_NCCL(g_nccl.AllReduce(dev_ptr, dev_ptr, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, g_comm, stream()));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return NEPB_OK;
}

struct nepb_contour {
    const nepb_spmf* op = nullptr;
    nepb_lu* lu = nullptr;
    int batch = 0, k = 0, mg = 0;
    DevBuf<double> vh, x, s, wgt, stage;
    std::vector<int> node_flags;
    int64_t nodes_done = 0;
};

int nepb_contour_create(const nepb_spmf* h, int k, int mg, int batch, nepb_contour** out) {
    NEPB_CHECK_ARG(h && out, "NULL argument");
    NEPB_CHECK_ARG(k >= 1 && k <= 256 && mg >= 1 && mg <= 64 && batch >= 1 && batch <= 4096, "bad sizes (k=%d mg=%d batch=%d)", k, mg, batch);
    *out = nullptr;
    nepb_contour* c = new nepb_contour();
    c->op = h;
    c->batch = batch;
    c->k = k;
    c->mg = mg;
    // dummy coefficients: the handle is (re)factorised per batch
    std::vector<double> coef((size_t)2 * batch * h->p, 0.0);
    for (int b = 0; b < batch; ++b) coef[(size_t)2 * b * h->p] = 1.0;
    LuSymbolicDev* sd = nullptr;
    int rc = lu_symbolic_get(h, &sd);
    if (rc) { delete c; return rc; }
    nepb_lu* lu = new nepb_lu();
    lu->op = h;
    lu->sym = sd;
    lu->nb = lu->cap = batch;
    const size_t nk = (size_t)h->n * k;
    cudaError_t e = lu->fronts.alloc((size_t)2 * batch * sd->S.front_total);
    if (e == cudaSuccess) e = lu->piv.alloc((size_t)batch * h->n);
    if (e == cudaSuccess) e = lu->info.alloc(batch);
    if (e == cudaSuccess) e = lu->coef.alloc((size_t)2 * batch * h->p);
    if (e == cudaSuccess) e = c->vh.alloc(2 * nk);
    if (e == cudaSuccess) e = c->x.alloc(2 * nk * batch);
    if (e == cudaSuccess) e = c->s.alloc(2 * nk * mg);
    if (e == cudaSuccess) e = c->wgt.alloc((size_t)2 * batch * mg);
    if (e != cudaSuccess) {
        set_error("contour workspace (batch %d, %.1f MB of fronts per node) does not fit: %s", batch, sd->S.front_total * 16e-6, cudaGetErrorString(e));
        delete lu;
        delete c;
        return e == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;
    }
    c->lu = lu;
    *out = c;
    return NEPB_OK;
}

int nepb_contour_destroy(nepb_contour* c) {
    if (c) {
        delete c->lu;
        delete c;
    }
    return NEPB_OK;
}

// Device part of one integration: S_dev = sum over this rank's nodes; asynchronous except for the per-batch status fetch.
int nepb_contour_integrate_dev(nepb_contour* c, int nnodes, const double* coef, const double* weights, int reduce) {
    NEPB_CHECK_ARG(c && (nnodes == 0 || (coef && weights)) && nnodes >= 0, "bad arguments");
    const nepb_spmf* h = c->op;
    const size_t nk = (size_t)h->n * c->k;
    NEPB_CUDA(cudaMemsetAsync(c->s.p, 0, sizeof(double) * 2 * nk * c->mg, stream()));
    c->node_flags.assign(nnodes, 0);
    for (int i0 = 0; i0 < nnodes; i0 += c->batch) {
        const int nb = std::min(c->batch, nnodes - i0);
        int rc = lu_refactor(c->lu, nb, coef + (size_t)2 * i0 * h->p);
        if (rc) return rc;
        NEPB_CUDA(cudaMemcpyAsync(c->wgt.p, weights + (size_t)2 * i0 * c->mg, sizeof(double) * 2 * nb * c->mg, cudaMemcpyHostToDevice, stream()));
        rc = lu_solve_device(c->lu, 0, nb, c->k, (const double2*)c->vh.p, 0, (double2*)c->x.p);
        if (rc) return rc;
        NEPB_LAUNCH(contour_accumulate_kernel, (unsigned)((nk + 255) / 256), 256, 0, nk, nb, c->mg, (const double2*)c->x.p, nk,
                    (const double2*)c->wgt.p, (double2*)c->s.p);
        NEPB_LAUNCH_CHECK();
        rc = lu_fetch_info(c->lu);  // also fences the host weight / coefficient buffers of this batch
        if (rc) return rc;
        for (int b = 0; b < nb; ++b) c->node_flags[i0 + b] = c->lu->h_info[b].flags | (c->lu->h_info[b].nperturbed ? 4 : 0);
    }
    c->nodes_done += nnodes;
    if (reduce) {
        int rc = nepb_comm_allreduce_sum_dev(c->s.p, (int64_t)(2 * nk * c->mg));
        if (rc) return rc;
    }
    return NEPB_OK;
}

int nepb_contour_set_probe(nepb_contour* c, const double* Vh, int64_t ldv) {
    NEPB_CHECK_ARG(c && Vh && ldv >= c->op->n, "bad arguments");
    return upload_colmajor(c->op->n, c->k, Vh, ldv, c->stage, c->vh.p, c->k, 0);
}

int nepb_contour_get_moments(nepb_contour* c, double* S) {
    NEPB_CHECK_ARG(c && S, "bad arguments");
    const int64_t n = c->op->n;
    for (int j = 0; j < c->mg; ++j) {
        int rc = download_colmajor(n, c->k, c->s.p + (size_t)2 * j * n * c->k, c->k, 0, c->stage, S + (size_t)2 * j * n * c->k, n);
        if (rc) return rc;
    }
    return NEPB_OK;
}

// Host-facing one-shot: probe in, moments out (n x k x mg column-major); node_flags[nnodes] optional.
int nepb_contour_integrate(nepb_contour* c, int nnodes, const double* coef, const double* weights, const double* Vh, int64_t ldv,
                           int reduce, double* S, int* node_flags) {
    int rc = nepb_contour_set_probe(c, Vh, ldv);
    if (rc) return rc;
    rc = nepb_contour_integrate_dev(c, nnodes, coef, weights, reduce);
    if (rc) return rc;
    rc = nepb_contour_get_moments(c, S);
    if (rc) return rc;
    if (node_flags) memcpy(node_flags, c->node_flags.data(), sizeof(int) * nnodes);
    return NEPB_OK;
}

}  // extern "C"
