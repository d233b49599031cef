// fullhist.cu -- full-system pair histogram (full_pairs_histograms_coords,
// Extensions/pairs_histograms.pyx:289-335) over the element-sorted, k-d ordered atom store (layout.h).
//
// Pipeline of one launch (full_hist_launch):
//   block_bbox_kernel    one bounding box per 256-record block and per 32-record sub-block
//   pair_list_kernel x2  + scan2_kernel: the (I tile, J block) pairs whose boxes are within maxDistance,
//                        cut into items of <= 8 surviving blocks (exact: a skipped pair cannot hold a hit)
//   full_hist_kernel     persistent CTAs take items from an atomic counter; sweep with the reference's fp32
//                        operation order, hits queued per lane and binned by the whole warp, counts in
//                        CTA-private shared memory, flushed with 64-bit atomics per element pair
//
// Bound: FP32 instruction issue on the pairs that cannot be excluded (19 exact operations per distance on the
// orthorhombic fast path), plus the bin pass of the in-range pairs.  Not HBM (one pass over 20 B/atom) and
// not tensor cores: the minimum image needs exact fp32 wrap/compare sequences that have no GEMM form.
// DESIGN.md section 4.1 and profiles/r1_fullhist_ncu_summary.md have the instruction budget and the ncu numbers.
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

// ------------------------------------------------------------------ the kernel
static const int JS = 512;   // J atoms staged per shared-memory sub-tile

// where an overflowing event goes when the reference's unchecked write is reproduced: straight to the
// global ordered histogram (rare: a pair within an ulp of maxDistance)
struct SpillTarget {
    unsigned long long *counts;   // [2][nEl*nEl][hs]
    long long cells;
    int slab_ab, slab_ba;
};

// Lane-private hit queues, warp-balanced binning.  In-range pairs are 0.3 % (plain sweep of a sparse box)
// to 50 % (maxDistance near half the box) of the pairs swept.  Binning a pair where it is found runs
// sqrt/div/atomic with a lane or two alive and, inlined per register atom and unroll step, bloats the loop
// past the instruction cache (ncu: "no instruction" was the top stall).  So the sweep only RECORDS a hit --
// a predicated 4-byte store of (i << 8 | q) into the lane's own column of a shared-memory array and a
// predicated pointer bump: no branch, no vote -- and when a column is nearly full or the staged block ends
// the warp bins ALL its columns together, entry f of the concatenated columns going to lane f % 32 (the
// owner column is found by a 5-step search over the scanned column lengths).  The I tile is kept in shared
// memory as well so that any lane can bin any entry.  The bin pass recomputes d2 from the same operands
// with the same instruction sequence, so it sees the identical value; integer counts make the order of
// binning irrelevant.
template <int R> struct SweepShape {
    static const int U = (R == 1) ? 4 : 2;          // J records per drain check
    static const int CAP = (R == 1) ? 20 : 16;      // queue entries per lane; [slot][thread] layout: own bank
};

template <int MODE, int R>
__device__ __forceinline__ void bin_warp_queues(const uint32_t *__restrict__ lq0 /* column of lane 0 of this warp */,
                                                int n, const float4 *__restrict__ sJ, const uint32_t *__restrict__ sO,
                                                const float4 *__restrict__ sI, const uint32_t *__restrict__ sIO,
                                                int i0, int jbase, bool tri, bool cross, const Lattice &L,
                                                const GridParams &g, unsigned int *__restrict__ sh,
                                                unsigned long long &ov, const SpillTarget &sp)
{
    const int lane = threadIdx.x & 31;
    int incl = n;                                  // inclusive scan of the column lengths
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - n;
    for (int f0 = 0; f0 < total; f0 += 32) {
        const int f = f0 + lane;
        int l = 0;                                 // owner column = number of columns that end at or before f
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, l + step - 1);
            if (v <= f) l += step;
        }
        const int first = __shfl_sync(0xffffffffu, excl, l & 31);
        if (f < total) {
            const uint32_t e = lq0[(f - first) * 256 + l];
            const int q = (int)(e & 0xFFu), i = (int)(e >> 8);
            if (!tri || (i0 + i < jbase + q)) {    // diagonal tile: only p < q counts
                const float4 a = sI[i], c = sJ[q];
                const float d2 = dist2<MODE>(a.x, a.y, a.z, c.x, c.y, c.z, L);
                const int b = bin_index(d2, g);
                const uint32_t mi = __float_as_uint(a.w), mj = __float_as_uint(c.w);
                // slot bit 1: inter-molecular, bit 0: the J atom comes first in original order
                const int slot = (((mi >> 8) == (mj >> 8)) ? 0 : 2) | ((cross && (sIO[i] > sO[q])) ? 1 : 0);
                if (b < g.hs) {
                    atomicAdd(&sh[slot * g.hs + b], 1u);
                } else {
                    ++ov;
                    const long long flat = (long long)((slot & 1) ? sp.slab_ba : sp.slab_ab) * g.hs + b;
                    if (g.spill && flat < sp.cells) atomicAdd(&sp.counts[((slot & 2) ? sp.cells : 0) + flat], 1ull);
                }
            }
        }
    }
    __syncwarp();
}

// one staged block of SEG_PAD J records against the thread's R register atoms
template <int MODE, int R>
__device__ __forceinline__ void sweep_block(const float4 *__restrict__ sJ, const uint32_t *__restrict__ sO, int jbase,
                                            const float (&xi)[R], const float (&yi)[R], const float (&zi)[R],
                                            const float4 *__restrict__ sI, const uint32_t *__restrict__ sIO, int i0,
                                            bool tri, bool cross, const Lattice &Lc, const GridParams &g, unsigned submask,
                                            uint32_t *__restrict__ lq, unsigned int *__restrict__ sh,
                                            unsigned long long &ov, const SpillTarget &sp)
{
    const int U = SweepShape<R>::U, CAP = SweepShape<R>::CAP;
    // lattice in plain registers: from the constant bank the compiler re-reads it (LDCU) every iteration
    Lattice L = Lc;
    if (MODE == MODE_ORTHO_FAST || MODE == MODE_ORTHO_GEN) {
        asm volatile("" : "+f"(L.b[0]), "+f"(L.b[4]), "+f"(L.b[8]));
    }
    float t2min = g.t2min, t2max = g.t2max;
    asm volatile("" : "+f"(t2min), "+f"(t2max));
    // 32-bit shared-window addresses: the push is STS + IADD under the hit predicate
    const uint32_t w0 = (uint32_t)__cvta_generic_to_shared(lq);
    const uint32_t wfull = w0 + (uint32_t)(CAP - U * R) * 1024u;
    uint32_t wp = w0;                            // next free slot of this lane's queue (stride 256 words)
    const uint32_t *lq0 = lq - (threadIdx.x & 31);
    uint32_t tid8 = threadIdx.x << 8;            // opaque, or the compiler rebuilds it under every hit predicate
    asm volatile("" : "+r"(tid8));
    // 32-record sub-blocks whose box is out of reach of every I block are not in submask (CTA-uniform)
    for (unsigned sm = submask; sm; sm &= sm - 1u) {
        const int qb = (__ffs(sm) - 1) * 32;
#pragma unroll 1
        for (int q = qb; q < qb + 32; q += U) {
            const uint32_t ev = tid8 + (uint32_t)q;          // entry = (r * SEG_PAD + tid) << 8 | (q + u)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float4 a = sJ[q + u];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float d2 = dist2<MODE>(xi[r], yi[r], zi[r], a.x, a.y, a.z, L);
                    if ((d2 >= t2min) && (d2 < t2max)) {
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(wp), "r"(ev + (uint32_t)(((r * SEG_PAD) << 8) + u)) : "memory");
                        wp += 1024u;
                    }
                }
            }
            if (__any_sync(0xffffffffu, wp > wfull)) {
                __syncwarp();
                bin_warp_queues<MODE, R>(lq0, (int)((wp - w0) >> 10), sJ, sO, sI, sIO, i0, jbase, tri, cross, Lc, g, sh, ov, sp);
                wp = w0;
            }
        }
    }
    __syncwarp();
    bin_warp_queues<MODE, R>(lq0, (int)((wp - w0) >> 10), sJ, sO, sI, sIO, i0, jbase, tri, cross, Lc, g, sh, ov, sp);
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
        if (lane == 0) {
            const size_t at = 2 * ((size_t)blk * (SEG_PAD / 32) + sb);
            sbox[at + 0] = make_float4(lo[0], lo[1], lo[2], 1e-6f * (1.0f + amax));
            sbox[at + 1] = make_float4(hi[0], hi[1], hi[2], (lo[0] <= hi[0]) ? 0.f : 1.f);
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
// then cut into ITEMS of at most PAIR_ITEM_BLOCKS surviving blocks, the units the sweep kernel's atomic
// counter hands out.  Items therefore cost about the same and none is empty: no CTA walks culled work, and
// the tail of a launch is one item long (what limited the 8-GPU efficiency of the chunked list).
// Two passes (count, exclusive scan, fill) size the lists exactly; the host reads the two totals back.
static const int PAIR_ITEM_BLOCKS = 8;

__global__ void pair_list_kernel(const WorkItem *__restrict__ rows, int n_rows, const float4 *__restrict__ bbox, CullParams cp,
                                 int *__restrict__ row_cnt, int *__restrict__ row_items, const int *__restrict__ row_start,
                                 const int *__restrict__ row_item_start, uint32_t *__restrict__ entries, int4 *__restrict__ items,
                                 int fill)
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

// counts layout (global, u64): [2][nEl*nEl][hs], index 0 = intra, 1 = inter.
// stats[0] += edge overflow events, stats[1] += (SEG_PAD I records x 32 J records) units actually swept.
template <int MODE, int R>
__global__ void __launch_bounds__(256, (R == 1) ? 4 : 3)
full_hist_kernel(const float4 *__restrict__ atoms, const uint32_t *__restrict__ orig, const float4 *__restrict__ bbox,
                 const WorkItem *__restrict__ rows, const uint32_t *__restrict__ entries, const int4 *__restrict__ items,
                 int n_items, int *__restrict__ next_item, Lattice L, GridParams g, CullParams cp, int nblocks, int nEl,
                 unsigned long long *__restrict__ counts, unsigned long long *__restrict__ stats)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sJ = reinterpret_cast<float4 *>(smem_raw);
    uint32_t *sO = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float4) * JS);
    unsigned char *cur = smem_raw + (sizeof(float4) + sizeof(uint32_t)) * JS;
    float4 *sI = reinterpret_cast<float4 *>(cur);                       cur += sizeof(float4) * SEG_PAD * R;
    uint32_t *sIO = reinterpret_cast<uint32_t *>(cur);                  cur += sizeof(uint32_t) * SEG_PAD * R;
    uint32_t *lq = reinterpret_cast<uint32_t *>(cur) + threadIdx.x;     cur += sizeof(uint32_t) * SweepShape<R>::CAP * 256;
    unsigned int *sh = reinterpret_cast<unsigned int *>(cur);
    __shared__ int s_item;

    const int tid = threadIdx.x, lane = threadIdx.x & 31;
    const int nsh = 4 * g.hs;
    for (int c = tid; c < nsh; c += 256) sh[c] = 0u;
    unsigned long long ov = 0, swept = 0;
    const long long cells = (long long)nEl * nEl * g.hs;
    int cur_slab = -1;                        // element pair the shared counters currently hold
    unsigned int since_flush = 0;             // block pairs swept into the counters since the last flush
    SpillTarget sp;
    sp.counts = counts; sp.cells = cells; sp.slab_ab = 0; sp.slab_ba = 0;

    while (true) {
        __syncthreads();                      // s_item consumed, previous item's sweeps done
        if (tid == 0) s_item = atomicAdd(next_item, 1);
        __syncthreads();
        const int it = s_item;
        WorkItem w;
        w.ea = -1; w.eb = -1;
        int4 item = make_int4(0, 0, 0, 0);    // {row, first entry, surviving J blocks (<= PAIR_ITEM_BLOCKS), -}
        if (it < n_items) { item = items[it]; w = rows[item.x]; }
        const int slab = (it < n_items) ? w.ea * nEl + w.eb : -2;
        // The counters belong to one element pair at a time; the item list is ordered by pair, so this
        // flush happens a handful of times per CTA (32-bit counters: also before 2^31 pairs pile up).
        if (slab != cur_slab || since_flush >= 32768u - 64u) {
            __syncthreads();                  // every warp has binned its queues (sweep_block ends with that)
            if (cur_slab >= 0) {
                for (int c = tid; c < nsh; c += 256) {
                    const unsigned int v = sh[c];
                    if (v) {
                        sh[c] = 0u;
                        const int slot = c / g.hs, b = c - slot * g.hs;
                        const long long at = ((slot >> 1) ? cells : 0) + (long long)((slot & 1) ? sp.slab_ba : sp.slab_ab) * g.hs + b;
                        atomicAdd(&counts[at], (unsigned long long)v);
                    }
                }
            }
            cur_slab = slab;
            since_flush = 0;
            sp.slab_ab = w.ea * nEl + w.eb; sp.slab_ba = w.eb * nEl + w.ea;
            __syncthreads();
        }
        if (it >= n_items) break;

        // the I tile: coordinates in registers for the sweep, the full records in shared memory for the
        // bin pass (made visible by the staging barrier below; the loop-top barrier protects the rewrite)
        float xi[R], yi[R], zi[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float4 a = make_float4(__int_as_float(0x7FC00000), __int_as_float(0x7FC00000), __int_as_float(0x7FC00000),
                                   __uint_as_float(PAD_META));     // NaN: never in range
            uint32_t o = 0xFFFFFFFFu;
            if (r < w.ni) {
                const int p = w.i0 + r * SEG_PAD + tid;
                a = atoms[p]; o = orig[p];
            }
            xi[r] = a.x; yi[r] = a.y; zi[r] = a.z;
            sI[r * SEG_PAD + tid] = a; sIO[r * SEG_PAD + tid] = o;
        }
        const bool cross = (w.ea != w.eb);
        const int bi = w.i0 / SEG_PAD;

        // the item's surviving J blocks, staged two at a time
        const bool same_el = (w.ea == w.eb);
        for (int e = 0; e < item.z; e += 2) {
            const int nb = (e + 1 < item.z) ? 2 : 1;
            const int jb0 = (int)entries[item.y + e] * SEG_PAD;
            const int jb1 = (nb == 2) ? (int)entries[item.y + e + 1] * SEG_PAD : jb0;
            __syncthreads();                  // previous staged blocks fully consumed
            sJ[tid] = atoms[jb0 + tid]; sO[tid] = orig[jb0 + tid];
            if (nb == 2) { sJ[SEG_PAD + tid] = atoms[jb1 + tid]; sO[SEG_PAD + tid] = orig[jb1 + tid]; }
            __syncthreads();
            // second level: lanes 0-7 / 8-15 test the 32-record sub-blocks of the two staged blocks (every warp
            // computes the same mask, so the CTA stays in step without exchanging it)
            unsigned sub = 0xFFFFu;
            if (cp.enabled) {
                const int which = lane >> 3, sb = lane & 7;
                bool nr = false;
                if (which < nb) {
                    const size_t at = 2 * ((size_t)nblocks + (size_t)((which ? jb1 : jb0) / SEG_PAD) * (SEG_PAD / 32) + sb);
                    const float4 loJ = bbox[at], hiJ = bbox[at + 1];
                    bool far = true;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r < w.ni) far = far && blocks_far(bbox[2 * (bi + r)], bbox[2 * (bi + r) + 1], loJ, hiJ, cp);
                    nr = !far;
                }
                sub = __ballot_sync(0xffffffffu, nr);
            }
            if (nb == 1) sub &= 0xFFu;
#pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                const int jbb = b ? jb1 : jb0;
                const bool tri = same_el && jbb < w.i0 + w.ni * SEG_PAD;      // J block inside the I tile: only p < q counts
                sweep_block<MODE, R>(sJ + b * SEG_PAD, sO + b * SEG_PAD, jbb, xi, yi, zi, sI, sIO, w.i0, tri, cross, L, g,
                                     (sub >> (8 * b)) & 0xFFu, lq, sh, ov, sp);
            }
            since_flush += (unsigned)(nb * w.ni);
            swept += (unsigned)(__popc(sub & 0xFFFFu) * w.ni);
        }
    }
    if (ov) atomicAdd(&stats[0], ov);
    if (tid == 0 && swept) atomicAdd(&stats[1], swept);
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

size_t full_hist_smem_bytes(int hs, int R)
{
    const size_t cap = (R == 1) ? SweepShape<1>::CAP : SweepShape<4>::CAP;
    return (sizeof(float4) + sizeof(uint32_t)) * (JS + (size_t)SEG_PAD * R) + sizeof(uint32_t) * cap * 256 +
           sizeof(unsigned int) * 4 * (size_t)hs;
}

template <int MODE, int R>
static int launch_full_t(cudaStream_t stream, int sm_count, const float4 *atoms, const uint32_t *orig, const float4 *bbox,
                         const WorkItem *rows, const uint32_t *entries, const int4 *items, int n_items, int *next_item,
                         const Lattice &L, const GridParams &g, const CullParams &cp, int nblocks, int nEl,
                         unsigned long long *counts, unsigned long long *stats)
{
    size_t smem = full_hist_smem_bytes(g.hs, R);
    FRMC_REQUIRE(smem <= 200 * 1024, FRMC_ELIMIT, "histSize %d needs %zu B of shared memory per CTA (limit 200 KiB)", g.hs, smem);
    auto kern = full_hist_kernel<MODE, R>;
    FRMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FRMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = std::min(n_items, sm_count * per_sm);
    if (grid < 1) return FRMC_OK;
    kern<<<grid, 256, smem, stream>>>(atoms, orig, bbox, rows, entries, items, n_items, next_item, L, g, cp, nblocks, nEl, counts, stats);
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
    cudaFree(row_ints); cudaFree(entries); cudaFree(items);
    row_ints = nullptr; entries = nullptr; items = nullptr; row_cap = entries_cap = items_cap = 0;
}

// Box pass and surviving-pair lists for `rows` under the culling parameters cp (bbox: 18 float4 per SEG_PAD
// block of scratch; lists: grow-only buffers).  Synchronises the stream once: the list sizes come back to
// the host.  Used by the histogram sweep below and by the distance-constraint sweep (atomdist.cu).
int build_pair_lists(cudaStream_t stream, const float4 *atoms, int64_t npad, float4 *bbox, const WorkItem *rows, int n_rows,
                     const CullParams &cp, PairLists &lists)
{
    const int nblocks = (int)(npad / SEG_PAD);
    lists.n_entries = lists.n_items = 0;
    if (n_rows <= 0 || nblocks <= 0) return FRMC_OK;
    if (cp.enabled) {
        block_bbox_kernel<<<(nblocks * 32 + 255) / 256, 256, 0, stream>>>(atoms, nblocks, cp.pbc, bbox);
        FRMC_LAUNCH_CHECK();
    }
    // row scratch: cnt, items, start, item_start [n_rows each] + 2 totals
    int rc = ensure_capacity(&lists.row_ints, &lists.row_cap, (size_t)4 * n_rows + 2, stream);
    if (rc) return rc;
    int *row_cnt = lists.row_ints, *row_items = row_cnt + n_rows, *row_start = row_items + n_rows,
        *row_item_start = row_start + n_rows, *totals = row_item_start + n_rows;
    const int list_grid = (int)(((long long)n_rows * 32 + 255) / 256);
    pair_list_kernel<<<list_grid, 256, 0, stream>>>(rows, n_rows, bbox, cp, row_cnt, row_items, nullptr, nullptr, nullptr, nullptr, 0);
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
                                                    lists.entries, lists.items, 1);
    FRMC_LAUNCH_CHECK();
    return FRMC_OK;
}

// Box pass, surviving-pair lists, then the sweep, on prepared device arrays.  next_item must be zeroed by the
// caller (stream-ordered) before every launch; stats[0] accumulates edge-overflow events, stats[1] swept
// (256 x 32) units.
int full_hist_launch(cudaStream_t stream, int sm_count, int mode, int R, const float4 *atoms, const uint32_t *orig,
                     int64_t npad, float4 *bbox, const WorkItem *rows, int n_rows, PairLists &lists, int *next_item,
                     const Lattice &L, const GridParams &g, int nEl, unsigned long long *counts, unsigned long long *stats)
{
    CullParams cp = make_cull(L, mode, g);
    if (g_no_cull) cp.enabled = 0;
    const int nblocks = (int)(npad / SEG_PAD);
    if (n_rows <= 0 || nblocks <= 0) return FRMC_OK;
    int rc = build_pair_lists(stream, atoms, npad, bbox, rows, n_rows, cp, lists);
    if (rc) return rc;
    if (lists.n_items == 0) return FRMC_OK;
#define FH_CASE(M)                                                                                         \
    case M:                                                                                                \
        return (R == 4) ? launch_full_t<M, 4>(stream, sm_count, atoms, orig, bbox, rows, lists.entries, lists.items, lists.n_items, next_item, L, g, cp, nblocks, nEl, counts, stats) \
                        : launch_full_t<M, 1>(stream, sm_count, atoms, orig, bbox, rows, lists.entries, lists.items, lists.n_items, next_item, L, g, cp, nblocks, nEl, counts, stats);
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

// Does block culling remove most of the sweep?  Estimated from the reach (maxDistance + one block
// extent, both ways) against the cell heights / coordinate spans; decides the register tiling below.
bool culling_pays(const Lattice &L, int mode, const float lo[3], const float hi[3], int64_t n, int nEl, const GridParams &g)
{
    if (g_no_cull || n <= 0) return false;
    const CullParams cp = make_cull(L, mode, g);
    if (!cp.enabled) return false;
    double span[3], vol = 1.0;
    for (int c = 0; c < 3; ++c) {
        span[c] = (mode == MODE_IBC) ? (double)hi[c] - (double)lo[c] : (double)cp.h[c];
        if (!(span[c] > 0.0) || !std::isfinite(span[c])) return false;
        vol *= span[c];
    }
    const double per_el = std::max(1.0, (double)n / std::max(1, nEl));
    const double extent = cbrt(vol * (double)SEG_PAD / per_el);
    const double reach = 2.0 * ((double)g.rmax + extent);
    double frac = 1.0;
    for (int c = 0; c < 3; ++c) frac *= std::min(1.0, reach / span[c]);
    return frac < 0.5;
}

// Tile-shape heuristic: returns R, the register atoms per thread (I-tile = 256*R).  R = 4 amortises the shared-
// memory load of a J record over four pairs (23.4 issue slots per pair instead of 26) and is right when
// every block pair has to be swept; when culling bites, R = 1 keeps the culled unit at one 256-atom block
// (about 2.5x fewer pairs swept than with a 1024-atom I-tile).
int choose_tiling(int64_t npad, bool sparse)
{
    return (npad >= 32768 && !sparse) ? 4 : 1;
}

int launch_counts64_to_float(cudaStream_t stream, const unsigned long long *counts, float *out, long long cells2)
{
    counts64_to_float_kernel<<<(unsigned)((cells2 + 255) / 256), 256, 0, stream>>>(counts, out, cells2);
    FRMC_LAUNCH_CHECK();
    return FRMC_OK;
}

}  // namespace frmc

using namespace frmc;

// Host-only inspection of the multi-GPU decomposition (no device needed): number of work items and
// of atom pairs covered by shard `shard` of `nshards` for a system with the given element indexes.
// Summed over the shards the pair count is n(n-1)/2; used by the CPU tests of the sharding logic.
extern "C" int frmc_set_block_culling(int on)
{
    int old = !g_no_cull;
    g_no_cull = on ? 0 : 1;
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
    const int R = choose_tiling(lay.npad, false);
    std::vector<WorkItem> items;
    build_rows(lay, R, shard, nshards, items);
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
    // the layout scratch lives for the thread: repeated calls (an Engine calling compute_data) reuse its pages, and
    // the two arrays that travel to the device are page-locked once per growth so the H2D copies run at link speed
    static thread_local HostLayout lay;
    static thread_local void *pinned_rec = nullptr, *pinned_orig = nullptr;
    {
        const size_t need = (size_t)n + (size_t)SEG_PAD * nEl;
        auto pin = [](auto &vec, size_t need_elems, void *&reg) {
            if (vec.capacity() >= need_elems && reg == (void *)vec.data()) return;
            if (reg) { cudaHostUnregister(reg); reg = nullptr; }
            if (vec.capacity() < need_elems) vec.reserve(need_elems + need_elems / 4);
            if (cudaHostRegister(vec.data(), vec.capacity() * sizeof(vec[0]), cudaHostRegisterPortable) == cudaSuccess) reg = vec.data();
            else cudaGetLastError();            // pageable copies still work
        };
        pin(lay.rec, need * 4, pinned_rec);
        pin(lay.orig, need, pinned_orig);
    }
    int rc = build_layout(coords, n, mol, el, nEl, isPBC, lay);
    if (rc) return rc;
    Lattice L;
    for (int i = 0; i < 9; ++i) L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    GridParams g = make_grid(rmin, rmax, bin, hs);
    int mode = choose_mode_from_bounds(L.b, isPBC, lay.lo, lay.hi);
    const int R = choose_tiling(lay.npad, culling_pays(L, mode, lay.lo, lay.hi, lay.n, nEl, g));
    std::vector<WorkItem> items;
    build_rows(lay, R, shard, nshards, items);

    float4 *d_atoms = (float4 *)ctx_buffer(c, 0, sizeof(float) * 4 * (size_t)lay.npad);
    uint32_t *d_orig = (uint32_t *)ctx_buffer(c, 1, sizeof(uint32_t) * (size_t)lay.npad);
    WorkItem *d_items = (WorkItem *)ctx_buffer(c, 2, sizeof(WorkItem) * items.size());
    unsigned long long *d_counts = (unsigned long long *)ctx_buffer(c, 4, sizeof(unsigned long long) * (2 * cells + 3));
    float4 *d_bbox = (float4 *)ctx_buffer(c, 3, sizeof(float4) * 18 * (size_t)(lay.npad / SEG_PAD + 1));
    float *d_out = (float *)ctx_buffer(c, 5, sizeof(float) * 2 * cells);
    if (!d_atoms || !d_orig || !d_items || !d_counts || !d_out || !d_bbox) return FRMC_ENOMEM;
    unsigned long long *d_ov = d_counts + 2 * cells;
    int *d_next = (int *)(d_counts + 2 * cells + 2);
    FRMC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * (2 * cells + 3), c->stream));
    if (lay.npad > 0) {
        FRMC_CUDA(cudaMemcpyAsync(d_atoms, lay.rec.data(), sizeof(float) * 4 * (size_t)lay.npad, cudaMemcpyHostToDevice, c->stream));
        FRMC_CUDA(cudaMemcpyAsync(d_orig, lay.orig.data(), sizeof(uint32_t) * (size_t)lay.npad, cudaMemcpyHostToDevice, c->stream));
    }
    if (!items.empty()) {
        FRMC_CUDA(cudaMemcpyAsync(d_items, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice, c->stream));
        static PairLists stateless_lists[64];          // per device, grow-only, like the context's scratch buffers
        rc = full_hist_launch(c->stream, c->sm_count, mode, R, d_atoms, d_orig, lay.npad, d_bbox, d_items, (int)items.size(),
                              stateless_lists[c->dev & 63], d_next, L, g, nEl, d_counts, d_ov);
        if (rc) return rc;
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
