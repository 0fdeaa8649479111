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
#include <map>
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

// One group = one stream + one batched factorisation workspace + its own moment accumulator.
struct ContourGroup {
    cudaStream_t st = nullptr;
    cudaStream_t side = nullptr;      // forward substitution runs here, beside the factorisation (lu_factor_solve_pipelined)
    std::vector<cudaEvent_t> ev;      // fork / per-level / join events of that pipeline
    cudaEvent_t done = nullptr;
    nepb_lu* lu = nullptr;
    DevBuf<double> x, s, wgt;
    int cap = 0;
    // pinned staging + one instantiated CUDA graph per batch share (the launch sequence depends only on the count)
    double* h_coef = nullptr;
    double* h_wgt = nullptr;
    LuInfo* h_info = nullptr;
    std::map<int, std::pair<cudaGraphExec_t, int>> graphs;  // count -> (exec, kernels in the graph)
    ~ContourGroup() {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second.first);
        if (h_coef) cudaFreeHost(h_coef);
        if (h_wgt) cudaFreeHost(h_wgt);
        if (h_info) cudaFreeHost(h_info);
        delete lu;
        if (done) cudaEventDestroy(done);
        for (auto e : ev) cudaEventDestroy(e);
        if (side) cudaStreamDestroy(side);
        if (st) cudaStreamDestroy(st);
    }
};

struct nepb_contour {
    const nepb_spmf* op = nullptr;
    int batch = 0, k = 0, mg = 0;
    std::vector<ContourGroup*> groups;
    DevBuf<double> vh, s, stage;
    DevBuf<const double*> d_parts;  // device copy of the group accumulator pointers
    std::vector<int> node_flags;
    int64_t nodes_done = 0;
    ~nepb_contour() {
        for (auto* g : groups) delete g;
    }
};

namespace nepb {
__global__ void __launch_bounds__(256) sum_groups_kernel(size_t count, int ng, const double* const* __restrict__ parts, double* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    double t = 0.0;
    for (int g = 0; g < ng; ++g) t += parts[g][idx];  // fixed order: reproducible
    out[idx] = t;
}
}  // namespace nepb

// (re)build the node groups of a contour handle on the symbolic analysis `sd`
static int contour_alloc_groups(nepb_contour* c, LuSymbolicDev* sd) {
    const nepb_spmf* h = c->op;
    for (auto* g : c->groups) delete g;
    c->groups.clear();
    // measured (profiles/r2_contour_breakdown.txt): 16 groups beat 8 by 1 % at 128 nodes and by 2-4 % at 16 nodes, plateau beyond;
    // the default stays 8: the group count fixes the order of the partial sums, and with 16 the rounding noise flips the order of a
    // conjugate eigenvalue pair of equal distance in test_beyn_dep0_disk_at_origin_with_sanity_check against the oracle's
    int ng = 8;
    if (const char* e = getenv("NEPB_CONTOUR_STREAMS")) ng = std::max(1, std::min(32, atoi(e)));
    ng = std::min(ng, c->batch);
    const int batch = c->batch, k = c->k, mg = c->mg;
    const size_t nk = (size_t)h->n * k;
    cudaError_t e = cudaSuccess;
    for (int g = 0; g < ng && e == cudaSuccess; ++g) {
        ContourGroup* G = new ContourGroup();
        c->groups.push_back(G);
        G->cap = (batch + ng - 1 - g) / ng;  // group sizes differ by at most one
        if (G->cap == 0) G->cap = 1;
        e = cudaStreamCreateWithFlags(&G->st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&G->done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&G->side, cudaStreamNonBlocking);
        for (int i = 0; i < 3 * sd->S.nlevels + 2 && e == cudaSuccess; ++i) {
            cudaEvent_t x = nullptr;
            e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
            if (e == cudaSuccess) G->ev.push_back(x);
        }
        nepb_lu* lu = new nepb_lu();
        G->lu = lu;
        lu->op = h;
        lu->sym = sd;
        lu->nb = lu->cap = G->cap;
        if (e == cudaSuccess) e = lu->fronts.alloc((size_t)2 * G->cap * sd->S.front_total);
        if (e == cudaSuccess) e = lu->piv.alloc((size_t)G->cap * h->n);
        if (e == cudaSuccess) e = lu->info.alloc(G->cap);
        if (e == cudaSuccess) e = lu->coef.alloc((size_t)2 * G->cap * h->p);
        if (e == cudaSuccess) e = G->x.alloc(2 * nk * G->cap);
        if (e == cudaSuccess) e = G->s.alloc(2 * nk * mg);
        if (e == cudaSuccess) e = G->wgt.alloc((size_t)2 * G->cap * mg);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&G->h_coef, sizeof(double) * 2 * G->cap * h->p);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&G->h_wgt, sizeof(double) * 2 * G->cap * mg);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&G->h_info, sizeof(LuInfo) * G->cap);
        if (e == cudaSuccess && lu_solve_reserve(lu, G->cap, k) != NEPB_OK) e = cudaErrorMemoryAllocation;
    }
    if (e != cudaSuccess) {
        set_error("contour workspace (batch %d, %.1f MB of fronts per node) does not fit: %s", batch, sd->S.front_total * 16e-6, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;
    }
    return NEPB_OK;
}

// `batch` nodes are in flight at a time, spread over NEPB_CONTOUR_STREAMS (default 8) groups; each group factorises and
// solves its share as one batched launch sequence on its own stream.
int nepb_contour_create(const nepb_spmf* h, int k, int mg, int batch, nepb_contour** out) {
    NEPB_CHECK_ARG(h && out, "NULL argument");
    NEPB_CHECK_ARG(k >= 1 && k <= 256 && mg >= 1 && mg <= 64 && batch >= 1 && batch <= 4096, "bad sizes (k=%d mg=%d batch=%d)", k, mg, batch);
    *out = nullptr;
    LuSymbolicDev* sd = nullptr;
    int rc = lu_symbolic_get(h, &sd);
    if (rc) return rc;
    nepb_contour* c = new nepb_contour();
    c->op = h;
    c->batch = batch;
    c->k = k;
    c->mg = mg;
    const size_t nk = (size_t)h->n * k;
    cudaError_t e = c->vh.alloc(2 * nk);
    if (e == cudaSuccess) e = c->s.alloc(2 * nk * mg);
    if (e != cudaSuccess) {
        set_error("contour moments do not fit: %s", cudaGetErrorString(e));
        delete c;
        return e == cudaErrorMemoryAllocation ? NEPB_E_NOMEM : NEPB_E_CUDA;
    }
    rc = contour_alloc_groups(c, sd);
    if (rc) {
        delete c;
        return rc;
    }
    *out = c;
    return NEPB_OK;
}

int nepb_contour_destroy(nepb_contour* c) {
    delete c;
    return NEPB_OK;
}

// The per-group pipeline for `cnt` nodes: coefficients / weights H2D (pinned), batched factorisation, batched solve with
// the shared probe, accumulate into the group's moments, status D2H.  Device work only -> captured once per count into a
// CUDA graph and replayed (one graph launch instead of ~400 kernel launches per group and step).
static int contour_group_enqueue(nepb_contour* c, ContourGroup* G, int cnt) {
    const nepb_spmf* h = c->op;
    const size_t nk = (size_t)h->n * c->k;
    G->lu->nb = cnt;
    NEPB_CUDA(cudaMemcpyAsync(G->lu->coef.p, G->h_coef, sizeof(double) * 2 * cnt * h->p, cudaMemcpyHostToDevice, G->st));
    NEPB_CUDA(cudaMemcpyAsync(G->wgt.p, G->h_wgt, sizeof(double) * 2 * cnt * c->mg, cudaMemcpyHostToDevice, G->st));
    static const bool pipelined = !(getenv("NEPB_CONTOUR_PIPELINE") && atoi(getenv("NEPB_CONTOUR_PIPELINE")) == 0);
    int rc;
    if (pipelined) {
        rc = lu_factor_solve_pipelined(G->lu, c->k, (const double2*)c->vh.p, 0, (double2*)G->x.p, G->side, G->ev.data());
        if (rc) return rc;
    } else {
        rc = lu_factor_device(G->lu);
        if (rc) return rc;
        rc = lu_solve_device(G->lu, 0, cnt, c->k, (const double2*)c->vh.p, 0, (double2*)G->x.p);
        if (rc) return rc;
    }
    NEPB_LAUNCH(contour_accumulate_kernel, (unsigned)((nk + 255) / 256), 256, 0, nk, cnt, c->mg, (const double2*)G->x.p, nk,
                (const double2*)G->wgt.p, (double2*)G->s.p);
    NEPB_LAUNCH_CHECK();
    NEPB_CUDA(cudaMemcpyAsync(G->h_info, G->lu->info.p, sizeof(LuInfo) * cnt, cudaMemcpyDeviceToHost, G->st));
    return NEPB_OK;
}

static int contour_group_run(nepb_contour* c, ContourGroup* G, int cnt) {
    static const bool use_graph = !(getenv("NEPB_CONTOUR_GRAPH") && atoi(getenv("NEPB_CONTOUR_GRAPH")) == 0);
    if (!use_graph) return contour_group_enqueue(c, G, cnt);
    auto it = G->graphs.find(cnt);
    if (it == G->graphs.end()) {
        const int64_t l0 = g_launches.load();
        cudaGraph_t graph = nullptr;
        NEPB_CUDA(cudaStreamBeginCapture(G->st, cudaStreamCaptureModeThreadLocal));
        int rc = contour_group_enqueue(c, G, cnt);
        cudaError_t e = cudaStreamEndCapture(G->st, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        NEPB_CUDA(e);
        const int nkern = (int)(g_launches.load() - l0);
        g_launches.fetch_sub(nkern);  // counted when the graph actually runs
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        NEPB_CUDA(e);
        it = G->graphs.emplace(cnt, std::make_pair(exec, nkern)).first;
    }
    NEPB_CUDA(cudaGraphLaunch(it->second.first, G->st));
    g_launches.fetch_add(it->second.second);
    return NEPB_OK;
}

// Device part of one integration: S_dev = sum over this rank's nodes (+ all-reduce).  Returns after the work finished.
static int contour_integrate_groups(nepb_contour* c, int nnodes, const double* coef, const double* weights) {
    const nepb_spmf* h = c->op;
    const size_t nk = (size_t)h->n * c->k;
    const int ng = (int)c->groups.size();
    // the probe upload (main stream) must be visible to every group stream
    NEPB_CUDA(cudaStreamSynchronize(main_stream()));
    for (auto* G : c->groups) NEPB_CUDA(cudaMemsetAsync(G->s.p, 0, sizeof(double) * 2 * nk * c->mg, G->st));
    c->node_flags.assign(nnodes, 0);
    int rc = NEPB_OK;
    for (int i0 = 0; i0 < nnodes && !rc; i0 += c->batch) {
        const int nb = std::min(c->batch, nnodes - i0);
        // contiguous share of the batch per group
        std::vector<int> first(ng + 1, 0);
        for (int g = 0; g < ng; ++g) first[g + 1] = first[g] + std::min(c->groups[g]->cap, std::max(0, (nb + ng - 1 - g) / ng));
        for (int g = 0; g < ng && !rc; ++g) {
            ContourGroup* G = c->groups[g];
            const int cnt = first[g + 1] - first[g];
            if (cnt <= 0) continue;
            const int j0 = i0 + first[g];
            set_current_stream(G->st);
            memcpy(G->h_coef, coef + (size_t)2 * j0 * h->p, sizeof(double) * 2 * cnt * h->p);
            memcpy(G->h_wgt, weights + (size_t)2 * j0 * c->mg, sizeof(double) * 2 * cnt * c->mg);
            rc = contour_group_run(c, G, cnt);
        }
        // status of this batch (also fences the host coefficient / weight buffers before the next batch reuses a group)
        for (int g = 0; g < ng && !rc; ++g) {
            ContourGroup* G = c->groups[g];
            const int cnt = first[g + 1] - first[g];
            if (cnt <= 0) continue;
            cudaError_t e = cudaStreamSynchronize(G->st);
            if (e != cudaSuccess) { set_error("CUDA error in the contour pipeline: %s", cudaGetErrorString(e)); rc = NEPB_E_CUDA; break; }
            for (int bidx = 0; bidx < cnt; ++bidx)
                c->node_flags[i0 + first[g] + bidx] = G->h_info[bidx].flags | (G->h_info[bidx].nperturbed ? 4 : 0) |
                                                        (lu_info_suspicious(G->h_info[bidx]) ? 16 : 0);
        }
    }
    reset_current_stream();
    if (rc) return rc;
    // S = sum of the group accumulators, on the main stream (all group streams are idle: lu_fetch_info synchronised them)
    std::vector<const double*> parts;
    for (auto* G : c->groups) parts.push_back(G->s.p);
    DevBuf<const double*>& d_parts = c->d_parts;  // per handle (and therefore per device)
    NEPB_CUDA(d_parts.reserve(parts.size()));
    NEPB_CUDA(cudaMemcpyAsync(d_parts.p, parts.data(), sizeof(double*) * parts.size(), cudaMemcpyHostToDevice, main_stream()));
    const size_t cnt = 2 * nk * c->mg;
    NEPB_LAUNCH(sum_groups_kernel, (unsigned)((cnt + 255) / 256), 256, 0, cnt, ng, (const double* const*)d_parts.p, c->s.p);
    NEPB_LAUNCH_CHECK();
    NEPB_CUDA(cudaStreamSynchronize(main_stream()));  // `parts` is a local buffer
    c->nodes_done += nnodes;
    return NEPB_OK;
}

int nepb_contour_integrate_dev(nepb_contour* c, int nnodes, const double* coef, const double* weights, int reduce) {
    NEPB_CHECK_ARG(c && (nnodes == 0 || (coef && weights)) && nnodes >= 0, "bad arguments");
    int rc = contour_integrate_groups(c, nnodes, coef, weights);
    reset_current_stream();
    if (rc) return rc;
    // static-pivoting fallback, as in nepb_lu_create: a node whose factorisation on the plain pattern met zero / tiny pivots
    // triggers one row matching (at that node), a rebuild of the groups on the matched analysis and a second pass
    static const int matching = getenv("NEPB_LU_MATCHING") ? atoi(getenv("NEPB_LU_MATCHING")) : 1;
    if (matching > 0 && !c->groups.empty() && !c->groups[0]->lu->sym->matched()) {
        int bad = -1;
        for (int i = 0; i < nnodes && bad < 0; ++i)
            if (c->node_flags[i] & (1 | 2 | 4 | 16)) bad = i;
        if (bad >= 0) {
            LuSymbolicDev* md = nullptr;
            rc = lu_symbolic_make_matched(c->op, coef + (size_t)2 * bad * c->op->p, &md);
            if (rc) return rc;
            rc = contour_alloc_groups(c, md);
            if (rc) return rc;
            rc = contour_integrate_groups(c, nnodes, coef, weights);
            reset_current_stream();
            if (rc) return rc;
        }
    }
    if (c->groups[0]->lu->sym->matched())
        for (int i = 0; i < nnodes; ++i) c->node_flags[i] |= 8;
    if (reduce) {
        rc = nepb_comm_allreduce_sum_dev(c->s.p, (int64_t)(2 * (size_t)c->op->n * c->k * c->mg));
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
    // reduce = 2: the extraction happens on rank 0 only (method_beyncontour.jl:114-184 runs once), the other ranks skip the download
    if (!(reduce == 2 && g_rank != 0)) {
        rc = nepb_contour_get_moments(c, S);
        if (rc) return rc;
    }
    if (node_flags) memcpy(node_flags, c->node_flags.data(), sizeof(int) * nnodes);
    return NEPB_OK;
}

}  // extern "C"
