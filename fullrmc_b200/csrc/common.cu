// common.cu -- host helpers: error string, device contexts, exact fp32 threshold search.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace frmc {

static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;
int g_edge_spill = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const char *last_error() { return g_err; }

// Smallest non-negative fp32 t with fl(sqrtf(t)) >= r.  Host sqrtf is IEEE correctly
// rounded (SSE sqrtss), identical to the device's __fsqrt_rn, and monotone, so the set
// {t : sqrtf(t) >= r} is an upper interval and the search below is exact.
float sqrt_threshold(float r)
{
    if (!(r > 0.0f)) return 0.0f;   // every distance is >= 0 >= r (also r = NaN: compare false -> never skipped)
    if (isinf(r)) return INFINITY;
    float t = r * r;
    if (isinf(t)) {
        // walk down from FLT_MAX; r is finite so sqrtf(FLT_MAX) may still be >= r
        t = 3.402823466e+38f;
        if (sqrtf(t) < r) return INFINITY;
    }
    volatile float s;
    // move up until the condition holds
    for (s = sqrtf(t); s < r; s = sqrtf(t)) t = nextafterf(t, INFINITY);
    // move down while the predecessor still satisfies it
    while (t > 0.0f) {
        float p = nextafterf(t, -INFINITY);
        s = sqrtf(p);
        if (s >= r) t = p; else break;
    }
    return t;
}

int choose_mode_from_bounds(const float *basis, int isPBC, const float lo[3], const float hi[3])
{
    if (!isPBC) return MODE_IBC;
    bool ortho = basis[1] == 0.0f && basis[2] == 0.0f && basis[3] == 0.0f && basis[5] == 0.0f &&
                 basis[6] == 0.0f && basis[7] == 0.0f;
    bool bounded = true;
    for (int c = 0; c < 3; ++c) {
        double span = (double)hi[c] - (double)lo[c];
        if (!(span < 1.49)) bounded = false;   // guarantees |fl(xi-xj)| < 1.5 (NaN spans fail too)
    }
    if (ortho) return bounded ? MODE_ORTHO_FAST : MODE_ORTHO_GEN;
    return bounded ? MODE_TRI_FAST : MODE_TRI_GEN;
}

int choose_mode(const float *basis, int isPBC, const float *coords, int64_t n)
{
    if (!isPBC) return MODE_IBC;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool finite = true;
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
            float v = coords[3 * i + c];
            if (!(v == v) || isinf(v)) finite = false;
            if (v < lo[c]) lo[c] = v;
            if (v > hi[c]) hi[c] = v;
        }
    if (n == 0) { lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.0f; }
    if (!finite) { hi[0] = INFINITY; }   // force the general wrap
    return choose_mode_from_bounds(basis, isPBC, lo, hi);
}

// ---------------------------------------------------------------- device contexts
static std::mutex g_ctx_mutex;
static DeviceCtx g_ctx[64];

DeviceCtx *get_ctx(int dev)
{
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("no CUDA device available (%s); fullrmc_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return nullptr;
    }
    if (dev < 0 || dev >= count || dev >= 64) {
        set_error("device index %d out of range (0..%d)", dev, count - 1);
        return nullptr;
    }
    DeviceCtx *c = &g_ctx[dev];
    if ((e = cudaSetDevice(dev)) != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(e));
        return nullptr;
    }
    if (c->dev < 0) {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
            return nullptr;
        }
        cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (c->sm_count <= 0) c->sm_count = 148;
        c->dev = dev;
    }
    return c;
}

void *ctx_buffer(DeviceCtx *c, int slot, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (c->cap[slot] >= bytes) return c->buf[slot];
    if (c->buf[slot]) { cudaStreamSynchronize(c->stream); cudaFree(c->buf[slot]); c->buf[slot] = nullptr; c->cap[slot] = 0; }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&c->buf[slot], want);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        c->buf[slot] = nullptr;
        return nullptr;
    }
    c->cap[slot] = want;
    return c->buf[slot];
}

void *ctx_pinned(DeviceCtx *c, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (c->pinned_cap >= bytes) return c->pinned;
    if (c->pinned) { cudaStreamSynchronize(c->stream); cudaFreeHost(c->pinned); c->pinned = nullptr; c->pinned_cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&c->pinned, want);
    if (e != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        c->pinned = nullptr;
        return nullptr;
    }
    c->pinned_cap = want;
    return c->pinned;
}

}  // namespace frmc

extern "C" {

const char *frmc_last_error(void) { return frmc::last_error(); }
const char *frmc_version(void) { return "fullrmc_b200 0.1 (sm_100a)"; }

int frmc_device_count(void)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) {
        frmc::set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
        return FRMC_ECUDA;
    }
    return count;
}

uint64_t frmc_launch_count(void) { return frmc::g_launch_count; }

int frmc_ctx_set_timing(int dev, int on)
{
    frmc::DeviceCtx *c = frmc::get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    if (on && !c->t0) {
        if (cudaEventCreate(&c->t0) != cudaSuccess || cudaEventCreate(&c->t1) != cudaSuccess) {
            frmc::set_error("cudaEventCreate failed");
            return FRMC_ECUDA;
        }
    }
    c->timing = on ? 1 : 0;
    c->timed = 0;
    return FRMC_OK;
}

int frmc_ctx_kernel_ms(int dev, double *ms)
{
    frmc::DeviceCtx *c = frmc::get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    if (!ms || !c->timed) { frmc::set_error("no timed kernel on device %d (frmc_ctx_set_timing first)", dev); return FRMC_ESTATE; }
    float f = 0.f;
    if (cudaEventSynchronize(c->t1) != cudaSuccess || cudaEventElapsedTime(&f, c->t0, c->t1) != cudaSuccess) {
        frmc::set_error("cudaEventElapsedTime failed");
        return FRMC_ECUDA;
    }
    *ms = (double)f;
    return FRMC_OK;
}

int frmc_set_edge_spill(int on)
{
    int old = frmc::g_edge_spill;
    frmc::g_edge_spill = on ? 1 : 0;
    return old;
}

}
