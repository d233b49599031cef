// common.cuh -- shared device/host helpers for the fullrmc_b200 CUDA library (sm_100a).
//
// The arithmetic in this file is the parity contract: every operation is an explicit
// round-to-nearest fp32 intrinsic (__fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn),
// which nvcc never contracts into FMA, in the reference's exact evaluation order
// (Extensions/pairs_distances.pyx:372-386, Extensions/pairs_histograms.pyx:58-63).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>

#include "../../include/fullrmc_b200.h"

namespace frmc {

// ---------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);
const char *last_error();                  // of the calling thread
extern unsigned long long g_launch_count;   // kernels launched by this library

#define FRMC_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            frmc::set_error("%s:%d %s failed: %s", __FILE__, __LINE__, #call,                \
                            cudaGetErrorString(_e));                                         \
            return FRMC_ECUDA;                                                               \
        }                                                                                    \
    } while (0)

#define FRMC_LAUNCH_CHECK()                                                                  \
    do {                                                                                     \
        ++frmc::g_launch_count;                                                              \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            frmc::set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,            \
                            cudaGetErrorString(_e));                                         \
            return FRMC_ECUDA;                                                               \
        }                                                                                    \
    } while (0)

#define FRMC_REQUIRE(cond, code, ...)                                                        \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            frmc::set_error(__VA_ARGS__);                                                    \
            return (code);                                                                   \
        }                                                                                    \
    } while (0)

// ---------------------------------------------------------------- geometry modes
// MODE selects the minimum-image variant at compile time so the inner loops carry no
// run-time branches.
enum : int {
    MODE_IBC = 0,        // infinite boundaries: plain Cartesian difference (pairs_distances.pyx:443-471)
    MODE_ORTHO_FAST = 1, // PBC, diagonal basis, all |frac diff| < 1.5  (sign-free wrap, 3 multiplies)
    MODE_TRI_FAST = 2,   // PBC, general basis,  all |frac diff| < 1.5
    MODE_ORTHO_GEN = 3,  // PBC, diagonal basis, unbounded fractional coordinates (floor/ceil wrap)
    MODE_TRI_GEN = 4     // PBC, general basis,  unbounded fractional coordinates
};

struct Lattice {
    float b[9];   // row-major basis, rows = lattice vectors (Engine.basisVectors)
};

// r-grid of one histogram + the d^2 thresholds equivalent to the reference's range test.
// Because IEEE sqrt is correctly rounded and monotone,
//   d = fl(sqrt(d2)) >= rmin  <=>  d2 >= t2min      and      d < rmax  <=>  d2 < t2max
// with t2min/t2max found on the host by exact fp32 search (see sqrt_threshold()).
struct GridParams {
    float rmin, rmax, bin;
    float t2min, t2max;
    int hs;
    int spill;      // 1: reproduce the reference's unchecked write when bin == histSize (flat index spills
    int pad;        //    into the next [a,b] slab of the same array); 0: drop.  Counted either way.
};

// global default for new grids / stateless calls (frmc_set_edge_spill)
extern int g_edge_spill;

// Lower bound on the reference's computed distance between any atom of block I and any atom of block J.
// Per axis the (periodic) separation of two points is at least |wrap(centre difference)| - half widths;
// the reference's per-axis round() wrap yields exactly that periodic separation.  Diagonal basis / no PBC:
// the axis bounds combine Euclidean-wise (h = |L_cc| or 1).  General basis: |r| >= |f_c| / |column c of
// B^-1| for every axis, so the largest single-axis bound is used (h_c = that reciprocal height).
// Everything errs on the near side: eps margins on the gaps, 1e-4 relative slack on the cut.
struct CullParams {
    float h[3];
    float t2cut;
    int pbc, euclid, enabled, pad;
};

// culling parameters for a geometry mode and the largest d^2 of interest (host; fullhist.cu)
CullParams make_cull(const Lattice &L, int mode, const GridParams &g);
extern int g_no_cull;   // debug: sweep every block pair (frmc_set_block_culling)

// smallest non-negative fp32 t such that fl(sqrtf(t)) >= r
float sqrt_threshold(float r);

// pick the geometry mode for a coordinate set (host side)
int choose_mode(const float *basis, int isPBC, const float *coords, int64_t n);
int choose_mode_from_bounds(const float *basis, int isPBC, const float lo[3], const float hi[3]);

// ---------------------------------------------------------------- exact fp32 device math
#ifdef __CUDACC__

// d - round(d) with round = half away from zero: floor(d+0.5) if d>0 else ceil(d-0.5)
// (pairs_distances.pyx:31-32).  Works for any finite d.
__device__ __forceinline__ float wrap_general(float d)
{
    float r = (d > 0.0f) ? floorf(__fadd_rn(d, 0.5f)) : ceilf(__fsub_rn(d, 0.5f));
    return __fsub_rn(d, r);
}

// Same value as wrap_general for |d| < 1.5: there fl(|d|+0.5) < 2, so the image is 1 exactly
// when fl(|d|+0.5) >= 1, i.e. |d| >= 0.5 - 2^-25 (0x3EFFFFFF; 0.5-2^-25+0.5 ties to even = 1.0),
// and 0 otherwise.  No FRND (quarter-rate conversion pipe) in the hot loop.
__device__ __forceinline__ float wrap_fast(float d)
{
    const float T = __int_as_float(0x3EFFFFFF);
    float one = __int_as_float((__float_as_int(d) & 0x80000000) | 0x3F800000);   // copysign(1, d)
    return (fabsf(d) >= T) ? __fsub_rn(d, one) : d;
}

// |wrap(d)| up to sign: enough for a diagonal basis, where only squares of the
// components reach the distance.  Saves the copysign.
__device__ __forceinline__ float wrap_fast_nosign(float d)
{
    // FSETP + predicated FADD (2 issue slots; the select form the compiler picks costs 3):
    // r = (|d| >= T) ? |d| - 1 : d     -- the sign of the unwrapped branch is irrelevant (squared later)
    float r;
    asm("{\n\t.reg .pred p;\n\t.reg .f32 a;\n\tabs.f32 a, %1;\n\tsetp.ge.f32 p, a, 0f3EFFFFFF;\n\t"
        "mov.f32 %0, %1;\n\t@p add.rn.f32 %0, a, 0fBF800000;\n\t}" : "=f"(r) : "f"(d));
    return r;
}

// squared real distance between two stored positions, reference operation order:
//   real_x = (bx*b00 + by*b10) + bz*b20 ... ; d2 = (rx*rx + ry*ry) + rz*rz
// (pairs_distances.pyx:380-386).  The result is symmetric in (i,j) because the wrap is odd.
template <int MODE>
__device__ __forceinline__ float dist2(float xi, float yi, float zi, float xj, float yj, float zj,
                                       const Lattice &L)
{
    float dx = __fsub_rn(xi, xj);
    float dy = __fsub_rn(yi, yj);
    float dz = __fsub_rn(zi, zj);
    float rx, ry, rz;
    if (MODE == MODE_IBC) {
        rx = dx; ry = dy; rz = dz;
    } else if (MODE == MODE_ORTHO_FAST || MODE == MODE_ORTHO_GEN) {
        // diagonal basis: by*b10 and bz*b20 are (signed) zeros, adding them is exact
        if (MODE == MODE_ORTHO_FAST) {
            dx = wrap_fast_nosign(dx); dy = wrap_fast_nosign(dy); dz = wrap_fast_nosign(dz);
        } else {
            dx = wrap_general(dx); dy = wrap_general(dy); dz = wrap_general(dz);
        }
        rx = __fmul_rn(dx, L.b[0]);
        ry = __fmul_rn(dy, L.b[4]);
        rz = __fmul_rn(dz, L.b[8]);
    } else {
        if (MODE == MODE_TRI_FAST) {
            dx = wrap_fast(dx); dy = wrap_fast(dy); dz = wrap_fast(dz);
        } else {
            dx = wrap_general(dx); dy = wrap_general(dy); dz = wrap_general(dz);
        }
        rx = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[0]), __fmul_rn(dy, L.b[3])), __fmul_rn(dz, L.b[6]));
        ry = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[1]), __fmul_rn(dy, L.b[4])), __fmul_rn(dz, L.b[7]));
        rz = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[2]), __fmul_rn(dy, L.b[5])), __fmul_rn(dz, L.b[8]));
    }
    return __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
}

// real-space difference vector point - other (PBC) with the reference's operation order
template <bool PBC>
__device__ __forceinline__ void diff3(float px, float py, float pz, float cx, float cy, float cz,
                                      const Lattice &L, float &rx, float &ry, float &rz)
{
    float dx = __fsub_rn(px, cx), dy = __fsub_rn(py, cy), dz = __fsub_rn(pz, cz);
    if (PBC) {
        dx = wrap_general(dx); dy = wrap_general(dy); dz = wrap_general(dz);
        rx = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[0]), __fmul_rn(dy, L.b[3])), __fmul_rn(dz, L.b[6]));
        ry = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[1]), __fmul_rn(dy, L.b[4])), __fmul_rn(dz, L.b[7]));
        rz = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[2]), __fmul_rn(dy, L.b[5])), __fmul_rn(dz, L.b[8]));
    } else {
        rx = dx; ry = dy; rz = dz;
    }
}

__device__ __forceinline__ bool blocks_far(const float4 loI, const float4 hiI, const float4 loJ, const float4 hiJ,
                                           const CullParams &cp)
{
    if (hiI.w == 1.f || hiJ.w == 1.f) return true;          // no finite atom on one side: nothing can be in range
    const float eps = loI.w + loJ.w;
    const float li[3] = {loI.x, loI.y, loI.z}, ui[3] = {hiI.x, hiI.y, hiI.z};
    const float lj[3] = {loJ.x, loJ.y, loJ.z}, uj[3] = {hiJ.x, hiJ.y, hiJ.z};
    float s = 0.f, m = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float half = 0.5f * ((ui[c] - li[c]) + (uj[c] - lj[c]));
        float d = 0.5f * ((lj[c] + uj[c]) - (li[c] + ui[c]));
        if (cp.pbc) d -= rintf(d);
        const float gap = fmaxf(fabsf(d) - half - eps, 0.f) * cp.h[c];
        s += gap * gap;
        m = fmaxf(m, gap * gap);
    }
    return (cp.euclid ? s : m) > cp.t2cut;
}

__device__ __forceinline__ bool in_range(float d2, const GridParams &g)
{
    return (d2 >= g.t2min) && (d2 < g.t2max);
}

// bin index of an in-range pair: (int)((d - rmin) / bin), fp32, truncation
// (pairs_histograms.pyx:63)
__device__ __forceinline__ int bin_index(float d2, const GridParams &g)
{
    float d = __fsqrt_rn(d2);
    return (int)__fdiv_rn(__fsub_rn(d, g.rmin), g.bin);
}

// bin rule applied to an already computed distance (the *_dists entry points)
__device__ __forceinline__ bool bin_of_distance(float d, const GridParams &g, int &b)
{
    if (d < g.rmin) return false;
    if (d >= g.rmax) return false;
    b = (int)__fdiv_rn(__fsub_rn(d, g.rmin), g.bin);
    return true;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------- per-device context
// One stream and a few grow-only scratch buffers per device, so the stateless entry
// points do not pay cudaMalloc/cudaFree on every call.
struct DeviceCtx {
    int dev = -1;
    cudaStream_t stream = nullptr;
    static const int NBUF = 24;
    void *buf[NBUF] = {nullptr};
    size_t cap[NBUF] = {0};
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    int sm_count = 0;
    // optional device timing of the stateless entry points' dominant kernel (frmc_ctx_set_timing): events around the
    // launch, on the context's stream; read back with frmc_ctx_kernel_ms after the call
    int timing = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int timed = 0;
};

struct CtxTimer {                 // RAII bracket: records t0 now and t1 when it goes out of scope
    DeviceCtx *c;
    explicit CtxTimer(DeviceCtx *ctx) : c(ctx) { if (c->timing && c->t0) { cudaEventRecord(c->t0, c->stream); } }
    ~CtxTimer() { if (c->timing && c->t1) { cudaEventRecord(c->t1, c->stream); c->timed = 1; } }
};

// returns nullptr (and sets the error string) on failure
DeviceCtx *get_ctx(int dev);
// grow-only device scratch buffer `slot`; returns nullptr on failure
void *ctx_buffer(DeviceCtx *c, int slot, size_t bytes);
void *ctx_pinned(DeviceCtx *c, size_t bytes);

}  // namespace frmc
