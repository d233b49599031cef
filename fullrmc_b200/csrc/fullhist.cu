// fullhist.cu -- full-system pair histogram (full_pairs_histograms_coords,
// Extensions/pairs_histograms.pyx:289-335) over the element-sorted, k-d ordered atom store (layout.h).
//
// Pipeline of one launch (full_hist_launch):
//   block_bbox_kernel    one bounding box per 256-record block and per 32-record sub-block
//   pair_list_kernel x2  + scan2_kernel: the (I block, J block) pairs whose boxes are within maxDistance,
//                        cut into items of <= 4 surviving blocks (exact: a skipped pair cannot hold a hit)
//   bin_table_kernel     the d^2 thresholds of every bin edge, found by exact search on the reference's own
//                        fp32 expression (int)((sqrt(d2) - rmin) / bin)
//   full_hist_warp_kernel  WARPS take (item, I sub-block) tasks from a per-element-pair atomic counter.  A warp
//                        holds 32 I atoms in registers, tests its own sub-block box against every 32-record J
//                        sub-block of the item, streams the survivors through a private two-stage shared-memory
//                        ring filled by TMA bulk copies (cp.async.bulk + mbarrier: no CTA barrier anywhere in the
//                        sweep), evaluates 32 x 32 distances per unit with the reference's fp32 operation order,
//                        queues (d2, j) of the hits per lane and bins them per sub-block into CTA-private
//                        shared-memory counters, flushed with 64-bit atomics when the CTA leaves an element pair
//
// Bound: FP32 instruction issue on the distance evaluations that cannot be excluded, plus the bin pass of the
// in-range pairs.  Not HBM (one pass over 20 B/atom) and not tensor cores: the minimum image needs exact fp32
// wrap/compare sequences that have no GEMM form.  DESIGN.md section 4.1 has the instruction budget.
#include "common.cuh"
#include "layout.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace frmc {

GridParams make_grid(float rmin, float rmax, float bin, int hs);
int g_no_cull = 0;
static int g_no_chunk_cull = 0;     // measurement: keep the unit-level culling only (frmc_set_chunk_culling)
int g_device_layout = 1;      // stateless full histogram: order the caller's atoms on the device (devlayout.cu)

// ------------------------------------------------------------------ host: layout + row list
// k-d ordering of one element's atoms: split the longest box axis at a record count that is a multiple
// of the unit (1024 = one register-tiled I-tile, then 256 = one block, then 32), so that every aligned
// group of 1024 / 256 / 32 consecutive records is a compact box -- about 2x fewer surviving block pairs
// than a Morton curve, whose 256-record runs are elongated.  nth_element on 16-byte points; the first
// levels fork threads (the stateless entry point pays this on every call).
struct KdPoint { float f[3]; uint32_t idx; };

// run fn(begin, end, part) over [0, n) in up to `parts` contiguous slices, one thread each
template <typename F>
static void parallel_slices(int64_t n, int parts, F fn)
{
    if (parts < 1) parts = 1;
    if (n < 65536 || parts == 1) { fn((int64_t)0, n, 0); return; }
    std::vector<std::thread> th;
    const int64_t step = (n + parts - 1) / parts;
    int part = 0;
    for (int64_t a = 0; a < n; a += step, ++part) th.emplace_back(fn, a, std::min(n, a + step), part);
    for (auto &t : th) t.join();
}

static void kd_order(KdPoint *a, size_t len, const float *lo, const float *hi, int fork_levels)
{
    size_t unit;
    if (len > 1024) unit = 1024; else if (len > 256) unit = 256; else if (len > 32) unit = 32; else return;
    const size_t nb = (len + unit - 1) / unit;
    const size_t k = (nb / 2) * unit;
    int ax = 0;
    if (hi[1] - lo[1] > hi[ax] - lo[ax]) ax = 1;
    if (hi[2] - lo[2] > hi[ax] - lo[ax]) ax = 2;
    if (ax == 0) std::nth_element(a, a + k, a + len, [](const KdPoint &x, const KdPoint &y) { return x.f[0] < y.f[0]; });
    else if (ax == 1) std::nth_element(a, a + k, a + len, [](const KdPoint &x, const KdPoint &y) { return x.f[1] < y.f[1]; });
    else std::nth_element(a, a + k, a + len, [](const KdPoint &x, const KdPoint &y) { return x.f[2] < y.f[2]; });
    const float cut = a[k].f[ax];
    float lhi[3] = {hi[0], hi[1], hi[2]}, rlo[3] = {lo[0], lo[1], lo[2]};
    lhi[ax] = cut; rlo[ax] = cut;
    if (fork_levels > 0 && len > 16384) {
        std::thread t(kd_order, a, k, lo, lhi, fork_levels - 1);
        kd_order(a + k, len - k, rlo, hi, fork_levels - 1);
        t.join();
    } else {
        kd_order(a, k, lo, lhi, 0);
        kd_order(a + k, len - k, rlo, hi, 0);
    }
}

int build_layout(const float *coords, int64_t n, const int32_t *mol, const int32_t *el, int nEl, int isPBC, HostLayout &out)
{
    FRMC_REQUIRE(n >= 0 && n < (1ll << 31) - 4096, FRMC_ELIMIT, "atom count %lld outside 0..2^31", (long long)n);
    FRMC_REQUIRE(nEl >= 1 && nEl <= FRMC_MAX_ELEMENTS, FRMC_ELIMIT, "numberOfElements %d outside 1..%d", nEl, FRMC_MAX_ELEMENTS);
    const bool timing = getenv("FRMC_LAYOUT_TIMING") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[layout] %-10s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
        tick = now;
    };
    out.n = n; out.nEl = nEl;
    out.seg_count.assign(nEl, 0);
    out.seg_start.assign(nEl + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        FRMC_REQUIRE(el[i] >= 0 && el[i] < nEl, FRMC_EINVAL, "elementIndex[%lld]=%d outside 0..%d", (long long)i, el[i], nEl - 1);
        out.seg_count[el[i]]++;
    }
    for (int e = 0; e < nEl; ++e) {
        int64_t padded = (out.seg_count[e] + SEG_PAD - 1) / SEG_PAD * SEG_PAD;
        out.seg_start[e + 1] = out.seg_start[e] + padded;
    }
    out.npad = out.seg_start[nEl];
    FRMC_REQUIRE(out.npad < (1ll << 31), FRMC_ELIMIT, "padded atom count exceeds 2^31");

    // molecule ids only matter through equality; use them directly when they fit 24 bits,
    // otherwise rank them (sort + unique)
    bool direct = true;
    for (int64_t i = 0; i < n; ++i)
        if (mol[i] < 0 || mol[i] >= 0x00FFFFFF) { direct = false; break; }
    std::vector<int32_t> rank;
    if (!direct) {
        std::vector<int32_t> keys(mol, mol + n);
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        FRMC_REQUIRE(keys.size() < 0x00FFFFFFu, FRMC_ELIMIT, "more than 2^24-1 distinct molecules");
        rank.resize(n);
        for (int64_t i = 0; i < n; ++i)
            rank[i] = (int32_t)(std::lower_bound(keys.begin(), keys.end(), mol[i]) - keys.begin());
    }

    {
        // largest spread of a molecule in original indexes: the sweep treats pairs farther apart than this as
        // inter-molecular without looking at their molecule indexes
        int32_t max_id = -1;
        for (int64_t i = 0; i < n; ++i) { const int32_t id = direct ? mol[i] : rank[i]; if (id > max_id) max_id = id; }
        std::vector<int32_t> first((size_t)max_id + 1, -1);
        int64_t span = 0;
        for (int64_t i = 0; i < n; ++i) {
            int32_t &f = first[(size_t)(direct ? mol[i] : rank[i])];
            if (f < 0) f = (int32_t)i; else span = std::max<int64_t>(span, i - f);
        }
        out.mol_span = (uint32_t)span;
    }
    lap("count+mol");
    out.rec.resize((size_t)out.npad * 4);
    out.orig.resize((size_t)out.npad);
    out.inv.resize((size_t)n);
    const float qnan = __builtin_nanf("");
    uint32_t padmeta = PAD_META;
    float padmeta_f;
    memcpy(&padmeta_f, &padmeta, 4);
    for (int e = 0; e < nEl; ++e)                         // only the padding records need the NaN fill
        for (int64_t p = out.seg_start[e] + out.seg_count[e]; p < out.seg_start[e + 1]; ++p) {
            out.rec[4 * p + 0] = qnan; out.rec[4 * p + 1] = qnan; out.rec[4 * p + 2] = qnan; out.rec[4 * p + 3] = padmeta_f;
            out.orig[p] = 0xFFFFFFFFu;
        }
    const int hw = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    {
        std::vector<float> plo((size_t)hw * 3, INFINITY), phi((size_t)hw * 3, -INFINITY);
        std::vector<int> pfin((size_t)hw, 1);
        parallel_slices(n, hw, [&](int64_t a, int64_t b, int part) {
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            int fin = 1;
            for (int64_t i = a; i < b; ++i)
                for (int c = 0; c < 3; ++c) {
                    const float v = coords[3 * i + c];
                    if (!(v == v) || isinf(v)) { fin = 0; continue; }
                    if (v < lo[c]) lo[c] = v;
                    if (v > hi[c]) hi[c] = v;
                }
            for (int c = 0; c < 3; ++c) { plo[(size_t)part * 3 + c] = lo[c]; phi[(size_t)part * 3 + c] = hi[c]; }
            pfin[(size_t)part] = fin;
        });
        out.finite = true;
        for (int c = 0; c < 3; ++c) { out.lo[c] = INFINITY; out.hi[c] = -INFINITY; }
        for (int t = 0; t < hw; ++t) {
            for (int c = 0; c < 3; ++c) {
                out.lo[c] = std::min(out.lo[c], plo[(size_t)t * 3 + c]);
                out.hi[c] = std::max(out.hi[c], phi[(size_t)t * 3 + c]);
            }
            if (!pfin[(size_t)t]) out.finite = false;
        }
    }

    lap("alloc+bounds");
    // gather each element's atoms (original order), then k-d order them on the periodically reduced
    // coordinates (a block must not straddle the seam because of integer offsets).  Slices of the input
    // are gathered in parallel: per-slice element counts give every slice its own write cursor per element.
    static_assert(sizeof(KdPoint) == 4 * sizeof(float), "KdPoint must overlay 4 floats");
    out.kd.resize((size_t)std::max<int64_t>(n, 1) * 4);
    KdPoint *pts = reinterpret_cast<KdPoint *>(out.kd.data());
    {
        std::vector<int64_t> slice_cnt((size_t)hw * nEl, 0);
        parallel_slices(n, hw, [&](int64_t a, int64_t b, int part) {
            int64_t *c = &slice_cnt[(size_t)part * nEl];
            for (int64_t i = a; i < b; ++i) c[el[i]]++;
        });
        std::vector<int64_t> cursor((size_t)hw * nEl, 0);
        int64_t at = 0;
        for (int e = 0; e < nEl; ++e)
            for (int t = 0; t < hw; ++t) { cursor[(size_t)t * nEl + e] = at; at += slice_cnt[(size_t)t * nEl + e]; }
        parallel_slices(n, hw, [&](int64_t a, int64_t b, int part) {
            int64_t *cur = &cursor[(size_t)part * nEl];
            for (int64_t i = a; i < b; ++i) {
                KdPoint &q = pts[(size_t)cur[el[i]]++];
                for (int c = 0; c < 3; ++c) {
                    const float v = coords[3 * i + c];
                    q.f[c] = ((v == v) && !isinf(v)) ? (isPBC ? (v - floorf(v)) : v) : 0.f;
                }
                q.idx = (uint32_t)i;
            }
        });
    }
    lap("gather");
    {
        // fork while a half still has > 16k points: short-lived threads, a few per core at most
        const int fork_levels = (hw > 1) ? 6 : 0;
        std::vector<std::thread> workers;
        int64_t at = 0;
        float blo[3], bhi[3];
        for (int c = 0; c < 3; ++c) {
            blo[c] = isPBC ? 0.f : out.lo[c];
            bhi[c] = isPBC ? 1.f : out.hi[c];
        }
        for (int e = 0; e < nEl; ++e) {
            KdPoint *base = pts + at;
            const size_t len = (size_t)out.seg_count[e];
            at += out.seg_count[e];
            if (len <= 32) continue;
            if (n > 65536 && hw > 1) workers.emplace_back(kd_order, base, len, blo, bhi, fork_levels);
            else kd_order(base, len, blo, bhi, 0);
        }
        for (auto &t : workers) t.join();
    }
    lap("kd");
    {
        // pts is element-major; record position = padded segment start + rank inside the element
        std::vector<int64_t> shift((size_t)nEl, 0);      // position - index into pts
        int64_t at = 0;
        for (int e = 0; e < nEl; ++e) { shift[e] = out.seg_start[e] - at; at += out.seg_count[e]; }
        parallel_slices(n, hw, [&](int64_t a, int64_t b, int) {
            for (int64_t k = a; k < b; ++k) {
                const int64_t i = pts[(size_t)k].idx;
                const int64_t p = k + shift[el[i]];
                uint32_t m = (uint32_t)(direct ? mol[i] : rank[i]);
                uint32_t meta = (m << 8) | (uint32_t)el[i];
                float mf;
                memcpy(&mf, &meta, 4);
                for (int c = 0; c < 3; ++c) out.rec[4 * p + c] = coords[3 * i + c];
                out.rec[4 * p + 3] = mf;
                out.orig[p] = (uint32_t)i;
                out.inv[i] = (int32_t)p;
            }
        });
    }
    lap("scatter");
    if (n == 0) for (int c = 0; c < 3; ++c) { out.lo[c] = 0.f; out.hi[c] = 0.f; }
    if (!out.finite) out.hi[0] = INFINITY;   // forces the general wrap
    return FRMC_OK;
}

void build_rows(const HostLayout &lay, int R, int shard, int nshards, std::vector<WorkItem> &rows)
{
    rows.clear();
    const int64_t TI = (int64_t)SEG_PAD * R;
    int64_t serial = 0;
    for (int ea = 0; ea < lay.nEl; ++ea) {
        if (lay.seg_count[ea] == 0) continue;
        const int64_t a0 = lay.seg_start[ea];
        const int64_t a1 = a0 + (lay.seg_count[ea] + SEG_PAD - 1) / SEG_PAD * SEG_PAD;
        for (int eb = ea; eb < lay.nEl; ++eb) {
            if (lay.seg_count[eb] == 0) continue;
            const int64_t b0 = lay.seg_start[eb];
            const int64_t b1 = b0 + (lay.seg_count[eb] + SEG_PAD - 1) / SEG_PAD * SEG_PAD;
            for (int64_t i0 = a0; i0 < a1; i0 += TI) {
                const int64_t i1 = std::min(i0 + TI, a1);
                // same element: q > p >= i0, records before the I-tile never pair with it
                const int64_t j0 = (ea == eb) ? i0 : b0;
                // rows are dealt out boustrophedon (0..S-1, S-1..0, ...): the J range of a same-element row
                // shrinks linearly with the tile index, and plain round-robin would hand shard 0 the larger row
                // of every group of S
                const int64_t pos = serial++ % (2 * (int64_t)nshards);
                if ((pos < nshards ? pos : 2 * (int64_t)nshards - 1 - pos) != shard) continue;
                WorkItem w;
                w.i0 = (int32_t)i0; w.ni = (int32_t)((i1 - i0) / SEG_PAD);
                w.j0 = (int32_t)j0; w.j1 = (int32_t)b1;
                w.ea = ea; w.eb = eb; w.pad0 = 0; w.pad1 = 0;
                rows.push_back(w);
            }
        }
    }
}

// ------------------------------------------------------------------ the sweep
// Work unit of the sweep: 32 I records (one per lane of a warp) x one 32-record J sub-block.
static const int V2_THREADS = 384;        // 12 warps per CTA share one set of counters (2 CTAs per SM at histSize 1000)
static const int V2_WARPS = V2_THREADS / 32;
static const int V2_ITEM_BLOCKS = 4;      // surviving J blocks per item: 8 tasks (one per I sub-block) of <= 32 units each
static const int V2_CAP = 23;             // hit-queue entries per lane (8 bytes each); emptied when a lane holds more than CAP - 8
static const int V2_TASK_LIMIT = 65536;   // tasks a CTA bins into its 32-bit counters between two flushes (32768 hits each at most)
// per-warp shared memory: queue [V2_CAP][32] uint2, J ring [2][32] float4, bin-pass scratch [32] uint4, 2 mbarriers
static const int V2_Q_BYTES = V2_CAP * 32 * 8;
static const int V2_SCRATCH_OFF = V2_Q_BYTES + 2 * 32 * 16;
static const int V2_MBAR_OFF = V2_SCRATCH_OFF + 32 * 16;
static const int V2_WARP_BYTES = V2_MBAR_OFF + 32;

__device__ __forceinline__ unsigned fh_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fh_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(fh_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fh_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fh_smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void fh_bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(fh_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(fh_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fh_mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok = 0, spins = 0;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(fh_smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins > 100000000u) __trap();     // never hang the GPU on a lost transaction
    } while (!ok);
}

// squared distance with the reference's operation order; NOWRAP drops the minimum-image step for units whose
// boxes prove |fl(xi - xj)| < 0.5 - 2^-25 on every axis: there round() is 0 and d - round(d) = d exactly
// (pairs_distances.pyx:31-32, 372-386), so the sequence of rounded operations is unchanged.
template <int MODE, bool NOWRAP>
__device__ __forceinline__ float dist2_unit(float xi, float yi, float zi, float xj, float yj, float zj, const Lattice &L)
{
    if (!NOWRAP || MODE == MODE_IBC) return dist2<MODE>(xi, yi, zi, xj, yj, zj, L);
    const float dx = __fsub_rn(xi, xj), dy = __fsub_rn(yi, yj), dz = __fsub_rn(zi, zj);
    float rx, ry, rz;
    if (MODE == MODE_ORTHO_FAST || MODE == MODE_ORTHO_GEN) {
        rx = __fmul_rn(dx, L.b[0]); ry = __fmul_rn(dy, L.b[4]); rz = __fmul_rn(dz, L.b[8]);
    } else {
        rx = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[0]), __fmul_rn(dy, L.b[3])), __fmul_rn(dz, L.b[6]));
        ry = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[1]), __fmul_rn(dy, L.b[4])), __fmul_rn(dz, L.b[7]));
        rz = __fadd_rn(__fadd_rn(__fmul_rn(dx, L.b[2]), __fmul_rn(dy, L.b[5])), __fmul_rn(dz, L.b[8]));
    }
    return __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
}

// The sweep reads SWEEP RECORDS {x, y, z, original index} (sweep_records_kernel: the store's records with the
// original index in place of the meta word).  A hit is queued as (d2, original index of the J atom) -- both already
// sit in registers, the d2 just computed and the last word of the J record -- so the push is a predicated 8-byte
// store and a predicated pointer bump, and the bin pass needs nothing else from the J side:
//   * ordered slot of a cross-element pair: original index of I against that of J;
//   * intra / inter: atoms of one molecule lie within `mol_span` original indexes of each other (the largest spread
//     of a molecule, found on the host), so a pair farther apart than that is inter-molecular without looking anything
//     up; the few candidates go through the exact comparison of the molecule indexes (slow path).
// Round 2, later: one warp per 32-record sub-block also RE-ORDERS its records -- the k-d order of the store stops at 32
// records; here it goes on to 16 and 8 (split along the longest axis of the run's box, full sort on (coordinate, lane),
// like the leaves of devlayout.cu) -- and writes the bounding boxes of the four 8-record CHUNKS of the sub-block
// (block_bbox_kernel's conventions).  The sweep tests a warp's I box against the chunk boxes of a J unit and skips
// the chunks that cannot hold a hit: at cfg5 that removes 30 % of the distance evaluations the unit boxes let through.
// The order inside a sub-block shows nowhere else: its own box is unchanged, both sides of the sweep read these
// records, the ordered [a,b] slot comes from the original indexes in the records.
__global__ void __launch_bounds__(256) sweep_records_kernel(const float4 *__restrict__ atoms, const uint32_t *__restrict__ orig, long long npad,
                                                            int pbc, float4 *__restrict__ out, float4 *__restrict__ cbox)
{
    __shared__ float4 s_rec[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * 8 + w;
    if (unit * 32 >= npad) return;
    const long long p = unit * 32 + lane;
    float4 a = atoms[p];
    a.w = __uint_as_float(orig[p]);
    for (int seg = 32; seg >= 16; seg >>= 1) {
        const bool fin = isfinite(a.x) && isfinite(a.y) && isfinite(a.z);
        float f[3] = {a.x, a.y, a.z};
        float lo[3], hi[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (pbc) f[c] = f[c] - floorf(f[c]);
            lo[c] = fin ? f[c] : INFINITY; hi[c] = fin ? f[c] : -INFINITY;
        }
        for (int o = seg >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
                hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
            }
        }
        int ax = 0;
        if (hi[1] - lo[1] > hi[ax] - lo[ax]) ax = 1;
        if (hi[2] - lo[2] > hi[ax] - lo[ax]) ax = 2;
        const float key = fin ? ((ax == 0) ? f[0] : (ax == 1) ? f[1] : f[2]) : INFINITY;     // records without a position go last
        const int base = lane & ~(seg - 1);
        int rank = 0;
        for (int j = 0; j < seg; ++j) {
            const float kj = __shfl_sync(0xffffffffu, key, base + j);
            rank += (kj < key || (kj == key && base + j < lane)) ? 1 : 0;
        }
        __syncwarp();
        s_rec[w][base + rank] = a;
        __syncwarp();
        a = s_rec[w][lane];
    }
    out[p] = a;
    // the chunk boxes: 8 consecutive records
    {
        const bool fin = isfinite(a.x) && isfinite(a.y) && isfinite(a.z);
        const float v[3] = {a.x, a.y, a.z};
        float lo[3], hi[3], amax = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f = pbc ? (v[c] - floorf(v[c])) : v[c];
            lo[c] = fin ? f : INFINITY; hi[c] = fin ? f : -INFINITY;
            if (fin) amax = fmaxf(amax, fabsf(v[c]));
        }
        for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
                hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
            }
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        }
        if ((lane & 7) == 0) {
            float4 *b = cbox + (size_t)unit * 8 + (size_t)(lane >> 3) * 2;
            b[0] = make_float4(lo[0], lo[1], lo[2], 1e-6f * (1.0f + amax));
            b[1] = make_float4(hi[0], hi[1], hi[2], (lo[0] <= hi[0]) ? 0.f : 1.f);
        }
    }
}

// 32-bit shared-window accesses: the bin pass addresses the queue and the bin-edge table with plain registers
__device__ __forceinline__ uint2 lds_u64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f64(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// what the rare events of the bin pass need; one copy per CTA in shared memory (the slab indexes change with the pair)
struct SlowCtx {
    GridParams g;
    unsigned long long *counts;      // [2][nEl*nEl][hs] global ordered histogram (edge spill)
    const int32_t *mol;              // molecule index by ORIGINAL atom index
    long long cells;
    int slab_ab, slab_ba;
    uint32_t sh_addr;                // shared address of the counters
    int cross;
    unsigned long long ov;           // edge-overflow events seen by this CTA
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct WarpCtx {                  // what the bin pass needs: warp-uniform or lane-private registers
    uint32_t wq;                  // shared address of the warp's queue ([V2_CAP][32] entries of 8 bytes; lane l owns column l)
    uint32_t w0;                  // wq + 8 * lane: this lane's column
    uint32_t sh_inter;            // shared address of the inter-molecular counters of the [a,b] slab
    uint32_t tab_addr;            // shared address of the bin-edge table
    uint32_t off_swap;            // byte offset from an [a,b] slab to its [b,a] slab (hs counters); 0 inside one element
    uint32_t oi;                  // original index of the lane's I atom
    uint32_t span;                // molecule spread in original indexes
    const SlowCtx *slow;          // shared memory
};

// The rare events, out of line: a pair that may be intra-molecular (exact molecule comparison), a bin guess the
// table did not confirm (exact fp32 sqrt and divide, the reference's own expression), an edge overflow (bin index ==
// histSize through fp32 rounding, or a grid whose maxDistance lies beyond rmin + hs * bin: counted, and written where
// the reference's unchecked store lands when spill emulation is on).
__device__ __noinline__ void bin_slow(float d2, uint32_t oi, uint32_t oj, SlowCtx *S)
{
    const GridParams g = S->g;
    const bool inter = (S->mol == nullptr) || S->mol[oi] != S->mol[oj];     // no array: every atom is its own molecule (mol_span == 0)
    const bool swp = S->cross && oi > oj;
    const int b = bin_index(d2, g);
    if ((unsigned)b < (unsigned)g.hs) {
        const uint32_t addr = S->sh_addr + 4u * (uint32_t)(((inter ? 2 : 0) + (swp ? 1 : 0)) * g.hs + b);
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
    } else {
        atomicAdd(&S->ov, 1ull);
        const long long flat = (long long)(swp ? S->slab_ba : S->slab_ab) * g.hs + b;
        if (g.spill && b >= 0 && flat < S->cells) atomicAdd(&S->counts[(inter ? S->cells : 0) + flat], 1ull);
    }
}

// The bin pass: the warp empties all 32 queue columns TOGETHER.  Hit counts differ a lot from lane to lane (an I atom
// on the near side of its sub-block sees several times the hits of one on the far side), so the concatenation of the
// columns is cut into 32 equal runs: lane l bins entries [l*c, (l+1)*c), c = ceil(total / 32), walking from its first
// (column, row) -- found by a 5-step search in the scanned column lengths -- across column ends.  What an entry needs
// from its I atom (the original index) comes from the column owner's slot in the scratch words.  Two entries per
// trip: the shared-memory loads of both (queue, bin edges) before the two counter updates.
// Fast path per entry: approximate sqrt -> bin guess -> the two d^2 edges of that bin confirm it (bin_table_kernel)
// -> one shared-memory increment.  Anything else goes through bin_slow().
template <bool TABLE>
__device__ __noinline__ uint32_t drain_queue(uint32_t wp, uint32_t wq, uint32_t oi, uint32_t off_swap, uint32_t sh_inter,
                                             uint32_t tab_addr, uint32_t span, int hs, float inv_bin, float c0, SlowCtx *S)
{
    const int lane = threadIdx.x & 31;
    const uint32_t w0 = wq + 8u * (uint32_t)lane;
    // scratch, one 16-byte record per NON-EMPTY column in lane order: {entries, original index of its I atom,
    // entries of all columns before it, shared address of the column}
    const uint32_t ws = wq + (uint32_t)V2_SCRATCH_OFF;
    const int n = (int)((wp - w0) >> 8);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return w0;
    const unsigned occupied = __ballot_sync(0xffffffffu, n > 0);
    const int n_cols = __popc(occupied);
    if (n > 0) {
        const uint32_t at = ws + 16u * (uint32_t)__popc(occupied & ((1u << lane) - 1u));
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(at), "r"((uint32_t)n), "r"(oi), "r"((uint32_t)(incl - n)), "r"(w0) : "memory");
    }
    __syncwarp();
    const int c = (total + 31) >> 5;
    const int f = lane * c;
    const int cnt = min(c, total - f);                            // <= 0: nothing left for this lane
    // the column holding entry f: the last record whose start (entries before it) is <= f
    int col = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int probe = col + step;
        if (probe < n_cols && (int)lds_u32(ws + 16u * (uint32_t)probe + 8u) <= f) col = probe;
    }
    uint4 rec;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rec.x), "=r"(rec.y), "=r"(rec.z), "=r"(rec.w) : "r"(ws + 16u * (uint32_t)col));
    int left = (int)rec.x - (f - (int)rec.z);                     // entries of this column from the lane's first one on
    uint32_t src = rec.w + 256u * (uint32_t)(f - (int)rec.z);     // shared address of the lane's next entry
    uint32_t oi_o = rec.y;
    const uint32_t span2 = 2u * span;
    for (int j = 0; j < c; j += 2) {
        float d2[2];
        uint32_t oj[2], oio[2];
        float2 t[2];
        int b[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            // every record holds at least one entry, so one step reaches the next entry (no loop, no branch)
            if (left == 0 && j + u < cnt) {
                ++col;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rec.x), "=r"(rec.y), "=r"(rec.z), "=r"(rec.w) : "r"(ws + 16u * (uint32_t)col));
                left = (int)rec.x; src = rec.w; oi_o = rec.y;
            }
            uint2 e = make_uint2(0u, 0u);
            if (j + u < cnt) e = lds_u64(src);
            src += 256u; --left;
            d2[u] = __uint_as_float(e.x); oj[u] = e.y; oio[u] = oi_o;
        }
        if (TABLE) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float s;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(d2[u]));
                b[u] = max(0, min(__float2int_rz(__fmaf_rn(s, inv_bin, c0)), hs - 1));
                t[u] = lds_f64(tab_addr + 8u * (uint32_t)b[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const bool maybe_intra = (oj[u] - oio[u] + span) <= span2;            // |oj - oi| <= span (unsigned wrap-around)
            const bool sure = TABLE && (d2[u] >= t[u].x) && (d2[u] < t[u].y) && !maybe_intra;
            if (j + u < cnt) {
                if (sure) {
                    const uint32_t addr = sh_inter + ((oio[u] > oj[u]) ? off_swap : 0u) + 4u * (uint32_t)b[u];
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
                } else {
                    bin_slow(d2[u], oio[u], oj[u], S);
                }
            }
        }
    }
    __syncwarp();
    return w0;
}

// one staged J sub-block (32 sweep records) against the lane's I atom.  TRI (the unit on the diagonal of the I block):
// only p < q counts, i.e. lane < record.
template <int MODE, bool NOWRAP, bool HASMIN, bool TABLE, bool TRI>
__device__ __forceinline__ void sweep_unit(const float4 *__restrict__ sJu, float xi, float yi, float zi, const Lattice &Lc,
                                           float t2min, float t2max, uint32_t &wp, const WarpCtx &W, int hs, float inv_bin,
                                           float c0, unsigned cmask)
{
    Lattice L = Lc;               // lattice in plain registers: from the constant bank the compiler re-reads it every iteration
    if (MODE == MODE_ORTHO_FAST || MODE == MODE_ORTHO_GEN) {
        asm volatile("" : "+f"(L.b[0]), "+f"(L.b[4]), "+f"(L.b[8]));
    }
    asm volatile("" : "+f"(t2min), "+f"(t2max));
    const uint32_t wfull = W.w0 + (uint32_t)(V2_CAP - 8) * 256u;
    if constexpr (TRI) {
        const int lane = threadIdx.x & 31;
#pragma unroll 1
        for (int q = 1; q < 32; ++q) {
            const float4 a = sJu[q];
            const float d2 = dist2_unit<MODE, NOWRAP>(xi, yi, zi, a.x, a.y, a.z, L);
            if ((!HASMIN || d2 >= t2min) && (d2 < t2max) && lane < q) {
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(wp), "r"(__float_as_uint(d2)), "r"(__float_as_uint(a.w)) : "memory");
                wp += 256u;
            }
            if ((q & 7) == 7 && __any_sync(0xffffffffu, wp > wfull)) { __syncwarp(); wp = drain_queue<TABLE>(wp, W.wq, W.oi, W.off_swap, W.sh_inter, W.tab_addr, W.span, hs, inv_bin, c0, (SlowCtx *)W.slow); }
        }
    } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            if (!((cmask >> c) & 1u)) continue;               // the chunk's box is out of reach of the warp's I atoms
            // eight records in flight: loads first, then eight independent distance chains, then the pushes
            float4 a[8];
            float d2[8];
            // (32-bit shared-window address in a register: a generic pointer costs a window-base computation per chunk)
            const uint32_t ja = fh_smem_u32(sJu) + 128u * (uint32_t)c;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w) : "r"(ja + 16u * (uint32_t)u));
#pragma unroll
            for (int u = 0; u < 8; ++u) d2[u] = dist2_unit<MODE, NOWRAP>(xi, yi, zi, a[u].x, a[u].y, a[u].z, L);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if ((!HASMIN || d2[u] >= t2min) && (d2[u] < t2max)) {
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(wp), "r"(__float_as_uint(d2[u])), "r"(__float_as_uint(a[u].w)) : "memory");
                    wp += 256u;
                }
            }
            // the queue holds V2_CAP entries per lane: it is emptied when a lane could not take eight more
            if (__any_sync(0xffffffffu, wp > wfull)) { __syncwarp(); wp = drain_queue<TABLE>(wp, W.wq, W.oi, W.off_swap, W.sh_inter, W.tab_addr, W.span, hs, inv_bin, c0, (SlowCtx *)W.slow); }
        }
    }
}

// ------------------------------------------------------------------ block bounding boxes + culling
// One record pair per SEG_PAD block, then one per 32-record sub-block: {lo.xyz, eps} {hi.xyz, empty}
// (18 float4 per block in all).  Under PBC the coordinates are first
// reduced to their fractional part (exact in fp32), so a block never straddles the periodic seam just
// because its atoms carry different integer offsets.  eps bounds the rounding of fl(xi - xj) on the raw
// coordinates.  Atoms with a non-finite component pair with nothing (their d2 is NaN/inf) and are left out.
__global__ void block_bbox_kernel(const float4 *__restrict__ atoms, int nblocks, int pbc, float4 *__restrict__ bbox)
{
    const int blk = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (blk >= nblocks) return;
    float4 *__restrict__ sbox = bbox + 2 * (size_t)nblocks;      // 8 sub-blocks of 32 records per block
    float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY}, bmax = 0.f;
    for (int sb = 0; sb < SEG_PAD / 32; ++sb) {
        const float4 a = atoms[(size_t)blk * SEG_PAD + sb * 32 + lane];
        const float v[3] = {a.x, a.y, a.z};
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, amax = 0.f;
        if (isfinite(a.x) && isfinite(a.y) && isfinite(a.z)) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                amax = fmaxf(amax, fabsf(v[c]));
                const float f = pbc ? (v[c] - floorf(v[c])) : v[c];
                lo[c] = f; hi[c] = f;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
                hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
            }
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        }
        // hi.w: 1 = no finite atom, 2 = every atom has all three coordinates in [0, 1) (raw = reduced: the sweep may
        // then prove from the boxes that no pair of two such sub-blocks needs the minimum-image step), 0 = otherwise
        const bool in_unit = !(isfinite(a.x) && isfinite(a.y) && isfinite(a.z)) ||
                             (a.x >= 0.f && a.x < 1.f && a.y >= 0.f && a.y < 1.f && a.z >= 0.f && a.z < 1.f);
        const bool all_unit = __all_sync(0xffffffffu, in_unit);
        if (lane == 0) {
            const size_t at = 2 * ((size_t)blk * (SEG_PAD / 32) + sb);
            sbox[at + 0] = make_float4(lo[0], lo[1], lo[2], 1e-6f * (1.0f + amax));
            sbox[at + 1] = make_float4(hi[0], hi[1], hi[2], (lo[0] <= hi[0]) ? (all_unit ? 2.f : 0.f) : 1.f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { blo[c] = fminf(blo[c], lo[c]); bhi[c] = fmaxf(bhi[c], hi[c]); }
        bmax = fmaxf(bmax, amax);
    }
    if (lane == 0) {
        bbox[2 * blk + 0] = make_float4(blo[0], blo[1], blo[2], 1e-6f * (1.0f + bmax));
        bbox[2 * blk + 1] = make_float4(bhi[0], bhi[1], bhi[2], (blo[0] <= bhi[0]) ? 0.f : 1.f);
    }
}

CullParams make_cull(const Lattice &L, int mode, const GridParams &g)
{
    CullParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.t2cut = g.t2max * 1.0001f;
    cp.enabled = std::isfinite(cp.t2cut) ? 1 : 0;
    cp.pbc = (mode != MODE_IBC);
    if (mode == MODE_IBC) {
        cp.euclid = 1; cp.h[0] = cp.h[1] = cp.h[2] = 1.0f;
    } else if (mode == MODE_ORTHO_FAST || mode == MODE_ORTHO_GEN) {
        cp.euclid = 1;
        cp.h[0] = fabsf(L.b[0]); cp.h[1] = fabsf(L.b[4]); cp.h[2] = fabsf(L.b[8]);
    } else {
        // heights of the cell: 1 / |column c of B^-1|, B rows = lattice vectors
        const double a[9] = {L.b[0], L.b[1], L.b[2], L.b[3], L.b[4], L.b[5], L.b[6], L.b[7], L.b[8]};
        const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
        if (!(fabs(det) > 0.0) || !std::isfinite(det)) { cp.enabled = 0; return cp; }
        double inv[9];
        inv[0] = (a[4] * a[8] - a[5] * a[7]) / det; inv[1] = (a[2] * a[7] - a[1] * a[8]) / det; inv[2] = (a[1] * a[5] - a[2] * a[4]) / det;
        inv[3] = (a[5] * a[6] - a[3] * a[8]) / det; inv[4] = (a[0] * a[8] - a[2] * a[6]) / det; inv[5] = (a[2] * a[3] - a[0] * a[5]) / det;
        inv[6] = (a[3] * a[7] - a[4] * a[6]) / det; inv[7] = (a[1] * a[6] - a[0] * a[7]) / det; inv[8] = (a[0] * a[4] - a[1] * a[3]) / det;
        cp.euclid = 0;
        for (int c = 0; c < 3; ++c) {
            const double col = sqrt(inv[c] * inv[c] + inv[3 + c] * inv[3 + c] + inv[6 + c] * inv[6 + c]);
            if (!(col > 0.0) || !std::isfinite(col)) { cp.enabled = 0; return cp; }
            cp.h[c] = (float)((1.0 / col) * (1.0 - 1e-5));
        }
    }
    for (int c = 0; c < 3; ++c) if (!std::isfinite(cp.h[c])) cp.enabled = 0;
    return cp;
}

// ------------------------------------------------------------------ surviving block pairs, built on the device
// A ROW (host, layout.h:WorkItem) is one I tile against the whole J range of one element pair.  One warp
// per row tests the row's J blocks 32 at a time against the tile's boxes and writes the survivors; rows are
// then cut into ITEMS of at most PairLists::item_blocks surviving blocks, the units the sweep kernels' atomic
// counters hand out.  Items therefore cost about the same and none is empty: no CTA walks culled work, and
// the tail of a launch is one item long (what limited the 8-GPU efficiency of the chunked list).
// Two passes (count, exclusive scan, fill) size the lists exactly; the host reads the two totals back.
__global__ void pair_list_kernel(const WorkItem *__restrict__ rows, int n_rows, const float4 *__restrict__ bbox, CullParams cp,
                                 int *__restrict__ row_cnt, int *__restrict__ row_items, const int *__restrict__ row_start,
                                 const int *__restrict__ row_item_start, uint32_t *__restrict__ entries, int4 *__restrict__ items,
                                 int fill, int PAIR_ITEM_BLOCKS)
{
    const int r = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const WorkItem w = rows[r];
    const int bi = w.i0 / SEG_PAD, jb0 = w.j0 / SEG_PAD, jb1 = w.j1 / SEG_PAD;
    const unsigned lt = (1u << lane) - 1u;
    const int base = fill ? row_start[r] : 0;
    int cnt = 0;
    for (int jr = jb0; jr < jb1; jr += 32) {
        const int bj = jr + lane;
        bool near = bj < jb1;
        if (near && cp.enabled) {
            const float4 loJ = bbox[2 * bj], hiJ = bbox[2 * bj + 1];
            bool far = true;
            for (int t = 0; t < w.ni; ++t) far = far && blocks_far(bbox[2 * (bi + t)], bbox[2 * (bi + t) + 1], loJ, hiJ, cp);
            near = !far;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, near);
        if (fill && near) entries[base + cnt + __popc(bal & lt)] = (uint32_t)bj;
        cnt += __popc(bal);
    }
    const int n_it = (cnt + PAIR_ITEM_BLOCKS - 1) / PAIR_ITEM_BLOCKS;
    if (!fill) {
        if (lane == 0) { row_cnt[r] = cnt; row_items[r] = n_it; }
    } else {
        const int ibase = row_item_start[r];
        for (int j = lane; j < n_it; j += 32)
            items[ibase + j] = make_int4(r, base + j * PAIR_ITEM_BLOCKS, min(PAIR_ITEM_BLOCKS, cnt - j * PAIR_ITEM_BLOCKS), 0);
    }
}

// exclusive scan of two int arrays of length n by ONE CTA of 1024 threads; totals[0], totals[1] receive the sums
__global__ void __launch_bounds__(1024) scan2_kernel(const int *__restrict__ a, const int *__restrict__ b, int n,
                                                     int *__restrict__ sa, int *__restrict__ sb, int *__restrict__ totals)
{
    __shared__ int pa[1024], pb[1024];
    const int t = threadIdx.x;
    const int chunk = (n + 1023) / 1024;
    const int lo = min(n, t * chunk), hi = min(n, lo + chunk);
    int xa = 0, xb = 0;
    for (int i = lo; i < hi; ++i) { xa += a[i]; xb += b[i]; }
    pa[t] = xa; pb[t] = xb;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int va = (t >= o) ? pa[t - o] : 0, vb = (t >= o) ? pb[t - o] : 0;
        __syncthreads();
        pa[t] += va; pb[t] += vb;
        __syncthreads();
    }
    int ra = pa[t] - xa, rb = pb[t] - xb;       // exclusive prefix of this thread's chunk
    for (int i = lo; i < hi; ++i) { sa[i] = ra; sb[i] = rb; ra += a[i]; rb += b[i]; }
    if (t == 1023) { totals[0] = pa[t]; totals[1] = pb[t]; }
}

// d^2 thresholds of the bin edges: T[b] = smallest fp32 d2 in [t2min, t2max) whose reference bin index
// (int)((sqrt(d2) - rmin) / bin) -- evaluated with the very IEEE operations of bin_index() -- is >= b, +inf when there
// is none; the expression is monotone in d2 (sqrt, subtraction, division by a positive bin and truncation all are),
// so a bisection over the ordered bit patterns of the positive floats finds it exactly.  tab[b] = (T[b], T[b+1]),
// b = 0..hs, with T[0] = -inf and T[hs+1] = +inf: the bin pass arrives at hs for the edge-overflow events.
__global__ void bin_table_kernel(GridParams g, float2 *__restrict__ tab)
{
    const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (b > g.hs) return;
    float edge[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int want = b + k;
        if (want == 0) { edge[k] = -INFINITY; continue; }
        if (want > g.hs) { edge[k] = INFINITY; continue; }
        unsigned lo = __float_as_uint(g.t2min), hi = __float_as_uint(g.t2max);      // search [lo, hi)
        const unsigned end = hi;
        while (lo < hi) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            if (bin_index(__uint_as_float(mid), g) >= want) hi = mid; else lo = mid + 1;
        }
        edge[k] = (lo < end) ? __uint_as_float(lo) : INFINITY;
    }
    tab[b] = make_float2(edge[0], edge[1]);
}

// everything one launch of the sweep needs, by value
struct SweepArgs {
    const float4 *recs;              // sweep records {x, y, z, original index}
    const float4 *cbox;              // per 32-record unit: {lo, hi} of its four 8-record chunks (sweep_records_kernel)
    const int32_t *mol;              // molecule index by original atom index (read only when mol_span > 0 ... or for candidates)
    const float4 *bbox;
    uint32_t mol_span;               // largest spread of a molecule in original indexes
    const WorkItem *rows; const int *pair_first_row; const int *row_item_start;
    const uint32_t *entries; const int4 *items; int *pair_next;
    const float2 *tab;
    unsigned long long *counts; unsigned long long *stats;
    int n_rows, n_pairs, n_items, nblocks, nEl;
    int chunk_cull;                  // 1: test the 8-record chunks of every unit (default); 0: unit-level culling only
    Lattice L; GridParams g; CullParams cp;
};

// counts layout (global, u64): [2][nEl*nEl][hs], index 0 = intra, 1 = inter.
// stats[0] += edge overflow events, stats[1] += (32 I records x 8 J records) chunks actually swept.
template <int MODE, bool HASMIN, bool TABLE>
__global__ void __launch_bounds__(V2_THREADS, 2) full_hist_warp_kernel(const SweepArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridParams &g = A.g;
    const int tid = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsh = 4 * g.hs;
    unsigned int *sh = reinterpret_cast<unsigned int *>(smem_raw);                      // 4 x hs counters of the CTA's element pair
    size_t off = ((size_t)nsh * 4 + 15) & ~(size_t)15;
    float2 *tab = reinterpret_cast<float2 *>(smem_raw + off);
    if (TABLE) off += (((size_t)g.hs + 1) * 8 + 15) & ~(size_t)15;
    unsigned char *wbase = smem_raw + off + (size_t)warp * V2_WARP_BYTES;
    float4 *sJ = reinterpret_cast<float4 *>(wbase + V2_Q_BYTES);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(wbase + V2_MBAR_OFF);
    __shared__ int s_tasks, s_again, s_p0;
    __shared__ int s_skip[2];
    __shared__ SlowCtx s_slow;

    for (int c = tid; c < nsh; c += V2_THREADS) sh[c] = 0u;
    if (TABLE) for (int c = tid; c <= g.hs; c += V2_THREADS) tab[c] = A.tab[c];
    if (lane == 0) {
        fh_mbar_init(&mbar[0], 1); fh_mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    const long long cells = (long long)A.nEl * A.nEl * g.hs;
    if (tid == 0) {
        s_tasks = 0; s_again = 0;
        s_slow.g = g; s_slow.counts = A.counts; s_slow.mol = A.mol; s_slow.cells = cells;
        s_slow.slab_ab = 0; s_slow.slab_ba = 0; s_slow.sh_addr = fh_smem_u32(sh); s_slow.cross = 0; s_slow.ov = 0ull;
        // CTAs start on the element pair that holds their share of the items, so that a CTA stays on one pair
        // (one flush of its counters) for most of the launch; pairs are then visited cyclically
        const long long target = ((2ll * blockIdx.x + 1) * A.n_items) / (2ll * gridDim.x);
        int p0 = 0;
        for (int p = 0; p < A.n_pairs; ++p) {
            const int r = A.pair_first_row[p];
            if ((r < A.n_rows ? A.row_item_start[r] : A.n_items) <= target) p0 = p;
        }
        s_p0 = p0;
    }
    __syncthreads();
    const int p0 = s_p0;

    WarpCtx W;
    W.wq = fh_smem_u32(wbase);
    W.w0 = W.wq + 8u * (uint32_t)lane;
    W.sh_inter = fh_smem_u32(sh) + 8u * (uint32_t)g.hs;
    W.tab_addr = fh_smem_u32(tab);
    W.off_swap = 0u; W.oi = 0u;
    W.span = (A.mol_span >= 0x40000000u) ? 0x7FFFFFFFu : A.mol_span;
    W.slow = &s_slow;
    const uint32_t sJ_addr = fh_smem_u32(sJ), mbar_addr = fh_smem_u32(mbar);
    uint32_t wp = W.w0;
    const float inv_bin = TABLE ? __frcp_rn(g.bin) : 0.f, c0 = TABLE ? -g.rmin * inv_bin : 0.f;
    unsigned long long swept = 0;
    int slab_ab = 0, slab_ba = 0;
    unsigned seq = 0;                             // J sub-blocks this warp has streamed: stage = seq & 1, parity = (seq >> 1) & 1
    const float T = __int_as_float(0x3EFFFFFF);   // 0.5 - 2^-25: below it round() is 0 (common.cuh:wrap_fast)

    for (int v = 0; v < A.n_pairs; ++v) {
        const int p = (p0 + v) % A.n_pairs;
        const int r0 = A.pair_first_row[p], r1 = A.pair_first_row[p + 1];
        const int is = (r0 < A.n_rows) ? A.row_item_start[r0] : A.n_items;
        const int ie = (r1 < A.n_rows) ? A.row_item_start[r1] : A.n_items;
        const int ntasks = (ie - is) * (SEG_PAD / 32);
        if (ntasks <= 0) continue;
        const WorkItem w0row = A.rows[r0];
        const bool cross = (w0row.ea != w0row.eb);
        slab_ab = w0row.ea * A.nEl + w0row.eb; slab_ba = w0row.eb * A.nEl + w0row.ea;
        W.off_swap = cross ? 4u * (uint32_t)g.hs : 0u;
        if (tid == 0) {
            s_slow.slab_ab = slab_ab; s_slow.slab_ba = slab_ba; s_slow.cross = cross ? 1 : 0;
            // a pair whose tasks have all been handed out is left at once (one barrier instead of three and a global atomic
            // per warp: a CTA that has run out of work walks through up to 14 such pairs at the end of the launch)
            s_skip[v & 1] = (*(volatile int *)&A.pair_next[p] >= ntasks) ? 1 : 0;
        }
        __syncthreads();                          // the slow path's view of the pair; the previous pair's flush is behind us
        if (s_skip[v & 1]) continue;
        bool again = true;
        while (again) {
            // ---- this warp's tasks of the pair
            while (true) {
                if (*(volatile int *)&s_tasks >= V2_TASK_LIMIT) { if (lane == 0) s_again = 1; break; }
                int t = 0;
                if (lane == 0) t = atomicAdd(&A.pair_next[p], 1);
                t = __shfl_sync(0xffffffffu, t, 0);
                if (t >= ntasks) break;
                if (lane == 0) atomicAdd(&s_tasks, 1);
                const int4 item = A.items[is + (t >> 3)];     // {row, first entry, surviving J blocks (<= V2_ITEM_BLOCKS), -}
                const int s = t & 7;                          // the I sub-block of the row's block this warp takes
                const WorkItem w = A.rows[item.x];
                const int bi = w.i0 / SEG_PAD;
                const size_t atI = 2 * ((size_t)A.nblocks + (size_t)bi * (SEG_PAD / 32) + s);
                const float4 loI = A.bbox[atI], hiI = A.bbox[atI + 1];
                if (A.cp.enabled && hiI.w == 1.f) continue;   // padding only
                const float4 ai = A.recs[w.i0 + s * 32 + lane];
                W.oi = __float_as_uint(ai.w);
                // which of the item's 32-record J sub-blocks can hold a hit for THIS warp's 32 atoms: lane = (entry, sub-block)
                const int e = lane >> 3, sb = lane & 7;
                int jb = 0;
                bool near = false, nowrap = false, tri = false;
                if (e < item.z) {
                    jb = (int)A.entries[item.y + e];
                    near = true;
                    if (!cross && jb == bi) { near = (sb >= s); tri = (sb == s); }     // inside the I block: only p < q counts
                    if (near && A.cp.enabled) {
                        const size_t atJ = 2 * ((size_t)A.nblocks + (size_t)jb * (SEG_PAD / 32) + sb);
                        const float4 loJ = A.bbox[atJ], hiJ = A.bbox[atJ + 1];
                        near = !blocks_far(loI, hiI, loJ, hiJ, A.cp);
                        nowrap = (hiI.w == 2.f) && (hiJ.w == 2.f) &&
                                 (__fsub_rn(hiI.x, loJ.x) < T) && (__fsub_rn(hiJ.x, loI.x) < T) &&
                                 (__fsub_rn(hiI.y, loJ.y) < T) && (__fsub_rn(hiJ.y, loI.y) < T) &&
                                 (__fsub_rn(hiI.z, loJ.z) < T) && (__fsub_rn(hiJ.z, loI.z) < T);
                    }
                }
                // which 8-record chunks of ITS unit the warp's I box can reach (sweep_records_kernel wrote their boxes): every
                // lane tests the four chunks of the unit it stands for; a unit none of whose chunks is in reach is dropped
                // (the diagonal unit is swept whole)
                unsigned cm = 0xFu;
                if (A.cp.enabled && A.chunk_cull && near && !tri) {
                    const float4 *cb = A.cbox + ((size_t)jb * (SEG_PAD / 32) + (size_t)sb) * 8;
                    float4 bx[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) bx[k] = __ldg(cb + k);
                    cm = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) cm |= blocks_far(loI, hiI, bx[2 * k], bx[2 * k + 1], A.cp) ? 0u : (1u << k);
                    near = cm != 0u;
                }
                unsigned m = __ballot_sync(0xffffffffu, near);
                const unsigned m_nowrap = __ballot_sync(0xffffffffu, near && nowrap);
                const unsigned m_tri = __ballot_sync(0xffffffffu, near && tri);
                if (!m) continue;
                // stream the surviving sub-blocks: unit n+1 is in flight (TMA) while unit n is swept.  A stage is
                // free as soon as its sweep has ended: queued hits carry everything the bin pass needs.
                auto issue = [&](int bit, unsigned sq) {
                    const int jbb = __shfl_sync(0xffffffffu, jb, bit & ~7);
                    if (lane == 0) {
                        // plain 32-bit shared-window addresses kept in registers: no generic->shared conversion per copy
                        const uint32_t st = sq & 1u, bar = mbar_addr + 8u * st, dst = sJ_addr + 512u * st;
                        const float4 *src = A.recs + ((size_t)jbb * SEG_PAD + (size_t)(bit & 7) * 32);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(512u) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                     ::"r"(dst), "l"(src), "r"(512u), "r"(bar) : "memory");
                    }
                };
                issue(__ffs(m) - 1, seq);
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1u;
                    __syncwarp();                              // every lane has read the stage the next copy overwrites
                    if (m) issue(__ffs(m) - 1, seq + 1u);
                    const int st = (int)(seq & 1u);
                    {
                        unsigned ok = 0, spins = 0;
                        const uint32_t bar = mbar_addr + 8u * (uint32_t)st, parity = (seq >> 1) & 1u;
                        do {
                            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                            if (!ok && ++spins > 100000000u) __trap();     // never hang the GPU on a lost transaction
                        } while (!ok);
                    }
                    const float4 *sJu = sJ + st * 32;
                    const unsigned cmask = __shfl_sync(0xffffffffu, cm, bit);      // the unit's chunks in reach
                    swept += (unsigned)__popc(cmask);
                    if ((m_tri >> bit) & 1u) {
                        if (MODE == MODE_IBC || ((m_nowrap >> bit) & 1u))
                            sweep_unit<MODE, true, HASMIN, TABLE, true>(sJu, ai.x, ai.y, ai.z, A.L, g.t2min, g.t2max, wp, W, g.hs, inv_bin, c0, cmask);
                        else
                            sweep_unit<MODE, false, HASMIN, TABLE, true>(sJu, ai.x, ai.y, ai.z, A.L, g.t2min, g.t2max, wp, W, g.hs, inv_bin, c0, cmask);
                    } else if (MODE == MODE_IBC || ((m_nowrap >> bit) & 1u)) {
                        sweep_unit<MODE, true, HASMIN, TABLE, false>(sJu, ai.x, ai.y, ai.z, A.L, g.t2min, g.t2max, wp, W, g.hs, inv_bin, c0, cmask);
                    } else {
                        sweep_unit<MODE, false, HASMIN, TABLE, false>(sJu, ai.x, ai.y, ai.z, A.L, g.t2min, g.t2max, wp, W, g.hs, inv_bin, c0, cmask);
                    }
                    ++seq;
                }
                __syncwarp();
                wp = drain_queue<TABLE>(wp, W.wq, W.oi, W.off_swap, W.sh_inter, W.tab_addr, W.span, g.hs, inv_bin, c0, &s_slow);   // the queue belongs to this task's I atoms
            }
            // ---- the CTA leaves the pair (or its 32-bit counters are due): counters -> global ordered histogram
            __syncthreads();
            const bool dirty = s_tasks > 0;
            again = s_again != 0;
            if (dirty) {
                for (int c = tid; c < nsh; c += V2_THREADS) {
                    const unsigned int cnt = sh[c];
                    if (cnt) {
                        sh[c] = 0u;
                        const int slot = c / g.hs, b = c - slot * g.hs;
                        const long long at = ((slot >> 1) ? cells : 0) + (long long)((slot & 1) ? slab_ba : slab_ab) * g.hs + b;
                        atomicAdd(&A.counts[at], (unsigned long long)cnt);
                    }
                }
            }
            __syncthreads();
            if (tid == 0) { s_tasks = 0; s_again = 0; }
            __syncthreads();
        }
    }
    if (tid == 0 && s_slow.ov) atomicAdd(&A.stats[0], s_slow.ov);      // the last flush's barriers are behind every bin pass
    if (lane == 0 && swept) atomicAdd(&A.stats[1], swept);
}

// counts (64-bit, SIGNED: the reference's running ordered arrays data-before+after may hold
// negative cells, because M removes a pair from [el_moved, el_other] while the full histogram
// had put it in [el_lower_index, el_higher_index]; only the symmetrised sum is a count)
// -> fp32 [nEl,nEl,hs] x2
__global__ void counts64_to_float_kernel(const unsigned long long *__restrict__ counts, float *__restrict__ out, long long cells2)
{
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells2) out[c] = (float)(long long)counts[c];
}

static size_t full_hist_smem_bytes(int hs, bool table)
{
    size_t s = (((size_t)4 * hs * 4 + 15) & ~(size_t)15) + (size_t)V2_WARPS * V2_WARP_BYTES;
    if (table) s += (((size_t)hs + 1) * 8 + 15) & ~(size_t)15;
    return s;
}

template <int MODE, bool HASMIN, bool TABLE>
static int launch_full_t(cudaStream_t stream, int sm_count, const SweepArgs &A)
{
    const size_t smem = full_hist_smem_bytes(A.g.hs, TABLE);
    FRMC_REQUIRE(smem <= 200 * 1024, FRMC_ELIMIT, "histSize %d needs %zu B of shared memory per CTA (limit 200 KiB)", A.g.hs, smem);
    auto kern = full_hist_warp_kernel<MODE, HASMIN, TABLE>;
    FRMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FRMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, V2_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long tasks = (long long)A.n_items * (SEG_PAD / 32);
    const int grid = (int)std::min<long long>((tasks + V2_WARPS - 1) / V2_WARPS, (long long)sm_count * per_sm);
    if (grid < 1) return FRMC_OK;
    kern<<<grid, V2_THREADS, smem, stream>>>(A);
    FRMC_LAUNCH_CHECK();
    return FRMC_OK;
}

// grow-only device buffer owned by the caller's PairLists
template <typename T>
static int ensure_capacity(T **buf, size_t *cap, size_t need, cudaStream_t stream)
{
    if (*cap >= need && *buf) return FRMC_OK;
    if (*buf) { FRMC_CUDA(cudaStreamSynchronize(stream)); cudaFree(*buf); *buf = nullptr; *cap = 0; }
    const size_t want = need + need / 4 + 1024;
    FRMC_CUDA(cudaMalloc((void **)buf, sizeof(T) * want));
    *cap = want;
    return FRMC_OK;
}

void PairLists::release()
{
    cudaFree(row_ints); cudaFree(entries); cudaFree(items); cudaFree(pair_next); cudaFree(bin_table); cudaFree(recs);
    row_ints = nullptr; entries = nullptr; items = nullptr; pair_next = nullptr; bin_table = nullptr; recs = nullptr;
    row_cap = entries_cap = items_cap = bin_cap = recs_cap = 0;
}

// Box pass and surviving-pair lists for `rows` under the culling parameters cp (bbox: 18 float4 per SEG_PAD
// block of scratch; lists: grow-only buffers).  Synchronises the stream once: the list sizes come back to
// the host.  Used by the histogram sweep below and by the distance-constraint sweep (atomdist.cu).
int build_pair_lists(cudaStream_t stream, const float4 *atoms, int64_t npad, float4 *bbox, const WorkItem *rows, int n_rows,
                     const CullParams &cp, PairLists &lists)
{
    const int nblocks = (int)(npad / SEG_PAD);
    lists.n_entries = lists.n_items = 0;
    lists.n_rows = n_rows;
    if (n_rows <= 0 || nblocks <= 0) return FRMC_OK;
    block_bbox_kernel<<<(nblocks * 32 + 255) / 256, 256, 0, stream>>>(atoms, nblocks, cp.pbc, bbox);
    FRMC_LAUNCH_CHECK();
    // row scratch: cnt, items, start, item_start [n_rows each] + 2 totals
    int rc = ensure_capacity(&lists.row_ints, &lists.row_cap, (size_t)4 * n_rows + 2, stream);
    if (rc) return rc;
    int *row_cnt = lists.row_ints, *row_items = row_cnt + n_rows, *row_start = row_items + n_rows,
        *row_item_start = row_start + n_rows, *totals = row_item_start + n_rows;
    const int list_grid = (int)(((long long)n_rows * 32 + 255) / 256);
    pair_list_kernel<<<list_grid, 256, 0, stream>>>(rows, n_rows, bbox, cp, row_cnt, row_items, nullptr, nullptr, nullptr, nullptr, 0, lists.item_blocks);
    FRMC_LAUNCH_CHECK();
    scan2_kernel<<<1, 1024, 0, stream>>>(row_cnt, row_items, n_rows, row_start, row_item_start, totals);
    FRMC_LAUNCH_CHECK();
    int h_tot[2] = {0, 0};
    FRMC_CUDA(cudaMemcpyAsync(h_tot, totals, sizeof(h_tot), cudaMemcpyDeviceToHost, stream));
    FRMC_CUDA(cudaStreamSynchronize(stream));
    FRMC_REQUIRE(h_tot[0] >= 0 && h_tot[1] >= 0, FRMC_ELIMIT, "more than 2^31 surviving block pairs");
    lists.n_entries = h_tot[0]; lists.n_items = h_tot[1];
    if (h_tot[1] == 0) return FRMC_OK;
    rc = ensure_capacity(&lists.entries, &lists.entries_cap, (size_t)h_tot[0], stream);
    if (rc) return rc;
    rc = ensure_capacity(&lists.items, &lists.items_cap, (size_t)h_tot[1], stream);
    if (rc) return rc;
    pair_list_kernel<<<list_grid, 256, 0, stream>>>(rows, n_rows, bbox, cp, row_cnt, row_items, row_start, row_item_start,
                                                    lists.entries, lists.items, 1, lists.item_blocks);
    FRMC_LAUNCH_CHECK();
    return FRMC_OK;
}

// Rows of one shard + the index of the first row of every element pair (rows are ordered by pair), packed into one
// blob for the device: [n_rows WorkItem][n_pairs + 1 int, the last = n_rows].
void pack_rows(const std::vector<WorkItem> &rows, std::vector<unsigned char> &blob, int &n_pairs)
{
    std::vector<int> first;
    for (size_t r = 0; r < rows.size(); ++r)
        if (r == 0 || rows[r].ea != rows[r - 1].ea || rows[r].eb != rows[r - 1].eb) first.push_back((int)r);
    n_pairs = (int)first.size();
    first.push_back((int)rows.size());
    blob.resize(rows.size() * sizeof(WorkItem) + first.size() * sizeof(int));
    if (!rows.empty()) memcpy(blob.data(), rows.data(), rows.size() * sizeof(WorkItem));
    memcpy(blob.data() + rows.size() * sizeof(WorkItem), first.data(), first.size() * sizeof(int));
}

// Box pass, surviving-pair lists, bin-edge table, then the sweep, on prepared device arrays.  `rows` is the device
// copy of a pack_rows() blob; mol_by_orig (device, molecule index by original atom index) is read only for pairs
// within mol_span original indexes of each other (HostLayout::mol_span).  stats[0] accumulates edge-overflow events,
// stats[1] swept (32 x 8) chunks.
int full_hist_launch(cudaStream_t stream, int sm_count, int mode, const float4 *atoms, const uint32_t *orig,
                     int64_t npad, float4 *bbox, const WorkItem *rows, int n_rows, int n_pairs, PairLists &lists,
                     const int32_t *mol_by_orig, uint32_t mol_span,
                     const Lattice &L, const GridParams &g, int nEl, unsigned long long *counts, unsigned long long *stats)
{
    CullParams cp = make_cull(L, mode, g);
    if (g_no_cull) cp.enabled = 0;
    const int nblocks = (int)(npad / SEG_PAD);
    if (n_rows <= 0 || nblocks <= 0) return FRMC_OK;
    lists.item_blocks = V2_ITEM_BLOCKS;
    int rc = build_pair_lists(stream, atoms, npad, bbox, rows, n_rows, cp, lists);
    if (rc) return rc;
    if (lists.n_items == 0) return FRMC_OK;
    const int max_pairs = FRMC_MAX_ELEMENTS * (FRMC_MAX_ELEMENTS + 1) / 2;
    FRMC_REQUIRE(n_pairs >= 1 && n_pairs <= max_pairs, FRMC_EINVAL, "bad element-pair count %d", n_pairs);
    if (!lists.pair_next) FRMC_CUDA(cudaMalloc((void **)&lists.pair_next, sizeof(int) * (max_pairs + 8)));
    FRMC_CUDA(cudaMemsetAsync(lists.pair_next, 0, sizeof(int) * (max_pairs + 8), stream));
    // the table replaces sqrt + divide in the bin pass when the grid is sane and it fits beside the counters
    const bool table = (g.bin > 0.f) && std::isfinite(g.bin) && std::isfinite(g.rmin) && g.hs <= 4096;
    if (table) {
        if (lists.bin_cap < (size_t)g.hs + 1) {
            if (lists.bin_table) { FRMC_CUDA(cudaStreamSynchronize(stream)); cudaFree(lists.bin_table); lists.bin_table = nullptr; }
            FRMC_CUDA(cudaMalloc((void **)&lists.bin_table, sizeof(float2) * ((size_t)g.hs + 1)));
            lists.bin_cap = (size_t)g.hs + 1;
        }
        bin_table_kernel<<<(g.hs + 1 + 127) / 128, 128, 0, stream>>>(g, lists.bin_table);
        FRMC_LAUNCH_CHECK();
    }
    // sweep records {x, y, z, original index} of the current coordinates
    if (lists.recs_cap < (size_t)npad) {
        if (lists.recs) { FRMC_CUDA(cudaStreamSynchronize(stream)); cudaFree(lists.recs); lists.recs = nullptr; }
        FRMC_CUDA(cudaMalloc((void **)&lists.recs, sizeof(float4) * ((size_t)npad + (size_t)npad / 4)));    // records + 8 box words per 32
        lists.recs_cap = (size_t)npad;
    }
    float4 *cbox = lists.recs + lists.recs_cap;
    sweep_records_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(atoms, orig, (long long)npad, cp.pbc, lists.recs, cbox);
    FRMC_LAUNCH_CHECK();
    SweepArgs A;
    A.recs = lists.recs; A.cbox = cbox; A.mol = mol_by_orig; A.mol_span = mol_span; A.bbox = bbox;
    A.rows = rows; A.pair_first_row = reinterpret_cast<const int *>(rows + n_rows); A.row_item_start = lists.row_ints + 3 * (size_t)n_rows;
    A.entries = lists.entries; A.items = lists.items; A.pair_next = lists.pair_next;
    A.tab = lists.bin_table; A.counts = counts; A.stats = stats;
    A.n_rows = n_rows; A.n_pairs = n_pairs; A.n_items = lists.n_items; A.nblocks = nblocks; A.nEl = nEl;
    A.L = L; A.g = g; A.cp = cp; A.chunk_cull = g_no_chunk_cull ? 0 : 1;
    const bool hasmin = g.t2min > 0.f;      // rmin <= 0: every non-NaN d2 passes the lower test
#define FH_CASE(M)                                                                                         \
    case M:                                                                                                \
        if (table) return hasmin ? launch_full_t<M, true, true>(stream, sm_count, A) : launch_full_t<M, false, true>(stream, sm_count, A); \
        return hasmin ? launch_full_t<M, true, false>(stream, sm_count, A) : launch_full_t<M, false, false>(stream, sm_count, A);
    switch (mode) {
        FH_CASE(MODE_IBC)
        FH_CASE(MODE_ORTHO_FAST)
        FH_CASE(MODE_TRI_FAST)
        FH_CASE(MODE_ORTHO_GEN)
        FH_CASE(MODE_TRI_GEN)
    }
#undef FH_CASE
    set_error("unknown geometry mode %d", mode);
    return FRMC_EINVAL;
}

int launch_counts64_to_float(cudaStream_t stream, const unsigned long long *counts, float *out, long long cells2)
{
    counts64_to_float_kernel<<<(unsigned)((cells2 + 255) / 256), 256, 0, stream>>>(counts, out, cells2);
    FRMC_LAUNCH_CHECK();
    return FRMC_OK;
}

}  // namespace frmc

using namespace frmc;

namespace frmc {
// per device, grow-only, like the context's scratch buffers
PairLists &stateless_lists_for(int dev)
{
    static PairLists lists[64];
    return lists[dev & 63];
}
}  // namespace frmc

// Host-only inspection of the multi-GPU decomposition (no device needed): number of work items and
// of atom pairs covered by shard `shard` of `nshards` for a system with the given element indexes.
// Summed over the shards the pair count is n(n-1)/2; used by the CPU tests of the sharding logic.
extern "C" int frmc_set_device_layout(int on)
{
    int old = g_device_layout;
    g_device_layout = on ? 1 : 0;
    return old;
}

extern "C" int frmc_set_block_culling(int on)
{
    int old = !g_no_cull;
    g_no_cull = on ? 0 : 1;
    return old;
}

extern "C" int frmc_set_chunk_culling(int on)
{
    int old = !g_no_chunk_cull;
    g_no_chunk_cull = on ? 0 : 1;
    return old;
}

// Host-only inspection of the store layout (no device needed): the original index of every record of the
// element-sorted, k-d ordered store (0xFFFFFFFF for padding) and the padded segment starts.  Used by the CPU
// tests of the ordering (permutation, element segments, compactness of the 256-record blocks).
extern "C" int frmc_debug_layout(int64_t n, const float *coords, const int32_t *mol, const int32_t *el, int nEl, int isPBC,
                                 int64_t capacity, uint32_t *orig_out, int64_t *npad_out, int64_t *seg_start_out)
{
    FRMC_REQUIRE(n >= 0 && coords && mol && el && orig_out && npad_out && seg_start_out, FRMC_EINVAL, "bad arguments");
    HostLayout lay;
    int rc = build_layout(coords, n, mol, el, nEl, isPBC, lay);
    if (rc) return rc;
    FRMC_REQUIRE(capacity >= lay.npad, FRMC_EINVAL, "orig_out holds %lld records, the layout has %lld", (long long)capacity, (long long)lay.npad);
    for (int64_t p = 0; p < lay.npad; ++p) orig_out[p] = lay.orig[(size_t)p];
    for (int e = 0; e <= nEl; ++e) seg_start_out[e] = lay.seg_start[(size_t)e];
    *npad_out = lay.npad;
    return FRMC_OK;
}

// The same inspection for the layout built on the device (devlayout.cu); needs a device.
extern "C" int frmc_debug_device_layout(int dev, int64_t n, const float *coords, const int32_t *mol, const int32_t *el, int nEl, int isPBC,
                                        int64_t capacity, uint32_t *orig_out, int64_t *npad_out, int64_t *seg_start_out)
{
    FRMC_REQUIRE(n >= 0 && coords && mol && el && orig_out && npad_out && seg_start_out, FRMC_EINVAL, "bad arguments");
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    HostLayout lay;
    float4 *d_atoms = nullptr;
    uint32_t *d_orig = nullptr;
    int32_t *d_keys = nullptr;
    int rc = device_layout(c, coords, n, mol, el, nEl, isPBC, lay, &d_atoms, &d_orig, &d_keys);
    if (rc) return rc;
    FRMC_REQUIRE(capacity >= lay.npad, FRMC_EINVAL, "orig_out holds %lld records, the layout has %lld", (long long)capacity, (long long)lay.npad);
    if (lay.npad > 0) FRMC_CUDA(cudaMemcpyAsync(orig_out, d_orig, sizeof(uint32_t) * (size_t)lay.npad, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    for (int e = 0; e <= nEl; ++e) seg_start_out[e] = lay.seg_start[(size_t)e];
    *npad_out = lay.npad;
    return FRMC_OK;
}

extern "C" int frmc_debug_work_items(int64_t n, const int32_t *el, int nEl, int shard, int nshards, int sm_count,
                                     int64_t *n_items, int64_t *n_pairs)
{
    FRMC_REQUIRE(n >= 0 && el && n_items && n_pairs, FRMC_EINVAL, "bad arguments");
    FRMC_REQUIRE(nshards >= 1 && shard >= 0 && shard < nshards, FRMC_EINVAL, "bad shard %d of %d", shard, nshards);
    std::vector<float> coords((size_t)n * 3, 0.f);
    std::vector<int32_t> mol((size_t)n, 0);
    HostLayout lay;
    int rc = build_layout(coords.data(), n, mol.data(), el, nEl, 1, lay);
    if (rc) return rc;
    (void)sm_count;
    std::vector<WorkItem> items;
    build_rows(lay, 1, shard, nshards, items);
    auto real_in = [&](int e, int64_t a, int64_t b) -> int64_t {   // real atoms of segment e inside positions [a, b)
        const int64_t end = lay.seg_start[e] + lay.seg_count[e];
        return std::max<int64_t>(0, std::min(b, end) - std::max(a, lay.seg_start[e]));
    };
    int64_t pairs = 0;
    for (const WorkItem &w : items) {
        const int64_t i0 = w.i0, i1 = w.i0 + (int64_t)w.ni * SEG_PAD;
        if (w.ea != w.eb) {
            pairs += real_in(w.ea, i0, i1) * real_in(w.eb, w.j0, w.j1);
        } else {
            const int64_t end = lay.seg_start[w.ea] + lay.seg_count[w.ea];
            for (int64_t p = i0; p < std::min(i1, end); ++p)                 // pairs p < q, q in [j0, j1)
                pairs += std::max<int64_t>(0, std::min<int64_t>(w.j1, end) - std::max<int64_t>(w.j0, p + 1));
        }
    }
    *n_items = (int64_t)items.size();
    *n_pairs = pairs;
    return FRMC_OK;
}

extern "C" int frmc_full_pairs_histograms_coords(int dev, const float *coords, int64_t n, const float *basis, int isPBC,
                                                 const int32_t *mol, const int32_t *el, int nEl, float rmin, float rmax,
                                                 float bin, int hs, int shard, int nshards, float *hintra,
                                                 float *hinter, uint64_t *edge_overflow)
{
    FRMC_REQUIRE(n >= 0, FRMC_EINVAL, "negative atom count");
    FRMC_REQUIRE(n == 0 || (coords && mol && el), FRMC_EINVAL, "NULL input array");
    FRMC_REQUIRE(hs >= 1 && hintra && hinter, FRMC_EINVAL, "bad histogram arguments");
    FRMC_REQUIRE(nshards >= 1 && shard >= 0 && shard < nshards, FRMC_EINVAL, "bad shard %d of %d", shard, nshards);
    FRMC_REQUIRE(nEl >= 1 && nEl <= FRMC_MAX_ELEMENTS, FRMC_ELIMIT, "numberOfElements %d outside 1..%d", nEl, FRMC_MAX_ELEMENTS);
    const int64_t cells = (int64_t)nEl * nEl * hs;
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    // the store layout of the caller's arrays: built on the device (devlayout.cu: raw arrays up, one small read-back);
    // frmc_set_device_layout(0) keeps the host k-d ordering of build_layout() instead (same histogram either way)
    static thread_local HostLayout lay;
    float4 *d_atoms = nullptr;
    uint32_t *d_orig = nullptr;
    int32_t *d_mol_dev = nullptr;
    int rc;
    if (g_device_layout) {
        rc = device_layout(c, coords, n, mol, el, nEl, isPBC, lay, &d_atoms, &d_orig, &d_mol_dev);
        if (rc) return rc;
    } else {
        rc = build_layout(coords, n, mol, el, nEl, isPBC, lay);
        if (rc) return rc;
        d_atoms = (float4 *)ctx_buffer(c, 0, sizeof(float) * 4 * (size_t)lay.npad);
        d_orig = (uint32_t *)ctx_buffer(c, 1, sizeof(uint32_t) * (size_t)lay.npad);
        if (!d_atoms || !d_orig) return FRMC_ENOMEM;
        if (lay.npad > 0) {
            FRMC_CUDA(cudaMemcpyAsync(d_atoms, lay.rec.data(), sizeof(float) * 4 * (size_t)lay.npad, cudaMemcpyHostToDevice, c->stream));
            FRMC_CUDA(cudaMemcpyAsync(d_orig, lay.orig.data(), sizeof(uint32_t) * (size_t)lay.npad, cudaMemcpyHostToDevice, c->stream));
        }
    }
    Lattice L;
    for (int i = 0; i < 9; ++i) L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    GridParams g = make_grid(rmin, rmax, bin, hs);
    int mode = choose_mode_from_bounds(L.b, isPBC, lay.lo, lay.hi);
    std::vector<WorkItem> items;
    build_rows(lay, 1, shard, nshards, items);
    std::vector<unsigned char> blob;
    int n_pairs = 0;
    pack_rows(items, blob, n_pairs);

    WorkItem *d_items = (WorkItem *)ctx_buffer(c, 2, blob.size());
    unsigned long long *d_counts = (unsigned long long *)ctx_buffer(c, 4, sizeof(unsigned long long) * (2 * cells + 3));
    float4 *d_bbox = (float4 *)ctx_buffer(c, 3, sizeof(float4) * 18 * (size_t)(lay.npad / SEG_PAD + 1));
    float *d_out = (float *)ctx_buffer(c, 5, sizeof(float) * 2 * cells);
    if (!d_items || !d_counts || !d_out || !d_bbox) return FRMC_ENOMEM;
    unsigned long long *d_ov = d_counts + 2 * cells;
    FRMC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * (2 * cells + 3), c->stream));
    if (!items.empty()) {
        FRMC_CUDA(cudaMemcpyAsync(d_items, blob.data(), blob.size(), cudaMemcpyHostToDevice, c->stream));
        int32_t *d_mol = d_mol_dev;                    // only molecular systems ever read it
        if (lay.mol_span > 0 && !d_mol) {
            d_mol = (int32_t *)ctx_buffer(c, 6, sizeof(int32_t) * (size_t)n);
            if (!d_mol) return FRMC_ENOMEM;
            FRMC_CUDA(cudaMemcpyAsync(d_mol, mol, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        }
        CtxTimer timer(c);                             // frmc_ctx_kernel_ms: box pass + lists + sweep
        rc = full_hist_launch(c->stream, c->sm_count, mode, d_atoms, d_orig, lay.npad, d_bbox, d_items, (int)items.size(), n_pairs,
                              stateless_lists_for(c->dev), d_mol, lay.mol_span, L, g, nEl, d_counts, d_ov);
        if (rc) return rc;
    }
    if (getenv("FRMC_LAYOUT_TIMING")) {
        static thread_local auto t_prev = std::chrono::steady_clock::now();
        cudaStreamSynchronize(c->stream);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[full histogram] lists+sweep done %.3f ms after the previous mark\n", std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    }
    rc = launch_counts64_to_float(c->stream, d_counts, d_out, 2 * cells);
    if (rc) return rc;
    unsigned long long ov = 0;
    FRMC_CUDA(cudaMemcpyAsync(hintra, d_out, sizeof(float) * cells, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(hinter, d_out + cells, sizeof(float) * cells, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(&ov, d_ov, sizeof(ov), cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    if (edge_overflow) *edge_overflow = ov;
    return FRMC_OK;
}
