// libnepb200 runtime: error strings, stream, event timer, device selection, MSWS stream.
#include "common.h"

#include <mutex>
#include <string.h>

namespace nepb {

static thread_local char t_err[1024] = "";
std::atomic<int64_t> g_launches{0};
static cudaStream_t g_stream = nullptr;
static bool g_stream_owned = false;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static int g_sms = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

static thread_local cudaStream_t t_override = nullptr;
static thread_local bool t_has_override = false;

// Work of the calling thread goes to `s` until reset_current_stream(): the contour loop spreads groups of quadrature
// nodes over several streams so that latency-bound levels of one group overlap throughput-bound levels of another.
void set_current_stream(cudaStream_t s) {
    t_override = s;
    t_has_override = true;
}
void reset_current_stream() { t_has_override = false; }
cudaStream_t main_stream();

cudaStream_t stream() {
    if (t_has_override) return t_override;
    return main_stream();
}

cudaStream_t main_stream() {
    if (!g_stream) {
        if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) == cudaSuccess) g_stream_owned = true;
    }
    return g_stream;
}

int sm_count() {
    if (g_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}

}  // namespace nepb

extern "C" {

const char* nepb_version(void) { return "nepb200 0.1 (sm_100a)"; }
const char* nepb_last_error(void) { return nepb::t_err; }

int nepb_device_count(int* count) {
    NEPB_CHECK_ARG(count, "count is NULL");
    *count = 0;
    NEPB_CUDA(cudaGetDeviceCount(count));
    return NEPB_OK;
}

int nepb_set_device(int device) {
    NEPB_CUDA(cudaSetDevice(device));
    nepb::g_sms = 0;
    return NEPB_OK;
}

int nepb_set_stream(void* s) {
    if (nepb::g_stream && nepb::g_stream_owned) cudaStreamDestroy(nepb::g_stream);
    nepb::g_stream = (cudaStream_t)s;
    nepb::g_stream_owned = false;
    if (!s) {
        NEPB_CUDA(cudaStreamCreateWithFlags(&nepb::g_stream, cudaStreamNonBlocking));
        nepb::g_stream_owned = true;
    }
    return NEPB_OK;
}

int nepb_synchronize(void) {
    NEPB_CUDA(cudaStreamSynchronize(nepb::stream()));
    return NEPB_OK;
}

int nepb_timer_start(void) {
    if (!nepb::g_ev0) {
        NEPB_CUDA(cudaEventCreate(&nepb::g_ev0));
        NEPB_CUDA(cudaEventCreate(&nepb::g_ev1));
    }
    NEPB_CUDA(cudaEventRecord(nepb::g_ev0, nepb::stream()));
    return NEPB_OK;
}

int nepb_timer_stop(float* ms) {
    NEPB_CHECK_ARG(ms && nepb::g_ev0, "timer not started");
    NEPB_CUDA(cudaEventRecord(nepb::g_ev1, nepb::stream()));
    NEPB_CUDA(cudaEventSynchronize(nepb::g_ev1));
    NEPB_CUDA(cudaEventElapsedTime(ms, nepb::g_ev0, nepb::g_ev1));
    return NEPB_OK;
}

int64_t nepb_launch_count(void) { return nepb::g_launches.load(); }

// ---- Middle-Square-Weyl-Sequence stream (host) -----------------------------------------------
// Same recurrence as src/gallery_extra/basic_random_examples.jl:73-95 (B. Widynski, arXiv 1704.00358),
// written with unsigned __int128.
typedef unsigned __int128 u128;

static inline u128 mk(uint64_t lo, uint64_t hi) { return ((u128)hi << 64) | lo; }

int nepb_msws_init(uint64_t seed_lo, uint64_t seed_hi, uint64_t st[6]) {
    NEPB_CHECK_ARG(st, "state is NULL");
    u128 base = mk(0xef01c4f2db0958c9ULL, 0x9ef09a97ac0f9ecaULL);
    u128 s = (mk(seed_lo, seed_hi) << 1) + base;
    u128 x = mk(0x3cbf13f7407cf43eULL, 0x1de568e1a1ca1b59ULL);
    u128 w = mk(0x5fafc1b7df9f9e0eULL, 0xd4ac5c288559e14aULL);
    st[0] = (uint64_t)x; st[1] = (uint64_t)(x >> 64);
    st[2] = (uint64_t)w; st[3] = (uint64_t)(w >> 64);
    st[4] = (uint64_t)s; st[5] = (uint64_t)(s >> 64);
    return NEPB_OK;
}

int nepb_msws_fill(uint64_t st[6], int64_t count, double* out) {
    NEPB_CHECK_ARG(st && (out || count == 0) && count >= 0, "bad arguments");
    u128 x = mk(st[0], st[1]), w = mk(st[2], st[3]), s = mk(st[4], st[5]);
    for (int64_t i = 0; i < count; ++i) {
        x *= x;
        w += s;
        x += w;
        x = (x >> 64) | (x << 64);
        out[i] = (double)(uint64_t)x / 18446744073709551616.0;
    }
    st[0] = (uint64_t)x; st[1] = (uint64_t)(x >> 64);
    st[2] = (uint64_t)w; st[3] = (uint64_t)(w >> 64);
    return NEPB_OK;
}

}  // extern "C"
