// fullhist.cu -- full-system pair histogram (full_pairs_histograms_coords,
// Extensions/pairs_histograms.pyx:289-335) as a tiled upper-triangle kernel over the
// element-sorted atom store (layout.h).
//
// Bound: FP32 issue slots (O(N^2) arithmetic on O(N) data that lives in shared memory
// and registers), plus shared-memory atomics for the in-range pairs.  Not HBM, not
// tensor cores: the minimum image needs exact fp32 wrap/compare sequences that have no
// GEMM form.  DESIGN.md section "Kernels" has the instruction budget.
#include "common.cuh"
#include "layout.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace frmc {

GridParams make_grid(float rmin, float rmax, float bin, int hs);

// ------------------------------------------------------------------ host: layout + work list
int build_layout(const float *coords, int64_t n, const int32_t *mol, const int32_t *el, int nEl, HostLayout &out)
{
    FRMC_REQUIRE(n >= 0 && n < (1ll << 31) - 4096, FRMC_ELIMIT, "atom count %lld outside 0..2^31", (long long)n);
    FRMC_REQUIRE(nEl >= 1 && nEl <= FRMC_MAX_ELEMENTS, FRMC_ELIMIT, "numberOfElements %d outside 1..%d", nEl, FRMC_MAX_ELEMENTS);
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

    out.rec.assign((size_t)out.npad * 4, 0.f);
    out.orig.assign((size_t)out.npad, 0xFFFFFFFFu);
    out.inv.assign((size_t)n, 0);
    const float qnan = __builtin_nanf("");
    uint32_t padmeta = PAD_META;
    float padmeta_f;
    memcpy(&padmeta_f, &padmeta, 4);
    for (int64_t p = 0; p < out.npad; ++p) {
        out.rec[4 * p + 0] = qnan; out.rec[4 * p + 1] = qnan; out.rec[4 * p + 2] = qnan; out.rec[4 * p + 3] = padmeta_f;
    }
    std::vector<int64_t> cursor(out.seg_start.begin(), out.seg_start.end() - 1);
    for (int c = 0; c < 3; ++c) { out.lo[c] = INFINITY; out.hi[c] = -INFINITY; }
    out.finite = true;
    for (int64_t i = 0; i < n; ++i) {
        int64_t p = cursor[el[i]]++;
        uint32_t m = (uint32_t)(direct ? mol[i] : rank[i]);
        uint32_t meta = (m << 8) | (uint32_t)el[i];
        float mf;
        memcpy(&mf, &meta, 4);
        for (int c = 0; c < 3; ++c) {
            float v = coords[3 * i + c];
            out.rec[4 * p + c] = v;
            if (!(v == v) || isinf(v)) out.finite = false;
            if (v < out.lo[c]) out.lo[c] = v;
            if (v > out.hi[c]) out.hi[c] = v;
        }
        out.rec[4 * p + 3] = mf;
        out.orig[p] = (uint32_t)i;
        out.inv[i] = (int32_t)p;
    }
    if (n == 0) for (int c = 0; c < 3; ++c) { out.lo[c] = 0.f; out.hi[c] = 0.f; }
    if (!out.finite) out.hi[0] = INFINITY;   // forces the general wrap
    return FRMC_OK;
}

void build_work_items(const HostLayout &lay, int R, int64_t chunkJ, int shard, int nshards, std::vector<WorkItem> &items)
{
    items.clear();
    const int64_t TI = (int64_t)SEG_PAD * R;
    int64_t serial = 0;
    for (int ea = 0; ea < lay.nEl; ++ea) {
        if (lay.seg_count[ea] == 0) continue;
        const int64_t a0 = lay.seg_start[ea];
        const int64_t a1 = a0 + (lay.seg_count[ea] + SEG_PAD - 1) / SEG_PAD * SEG_PAD;
        for (int64_t i0 = a0; i0 < a1; i0 += TI) {
            const int64_t i1 = std::min(i0 + TI, a1);
            for (int eb = ea; eb < lay.nEl; ++eb) {
                if (lay.seg_count[eb] == 0) continue;
                const int64_t b0 = lay.seg_start[eb];
                const int64_t b1 = b0 + (lay.seg_count[eb] + SEG_PAD - 1) / SEG_PAD * SEG_PAD;
                for (int64_t c0 = b0; c0 < b1; c0 += chunkJ) {
                    const int64_t j1 = std::min(c0 + chunkJ, b1);
                    int64_t j0 = c0;
                    int tri = 0;
                    if (ea == eb) {
                        if (j1 <= i0 + 1) continue;       // no q > p in this chunk
                        if (j0 < i0) j0 = i0;             // q > p >= i0: records before the I-tile never pair with it
                        tri = (j0 < i1) ? 1 : 0;          // ranges overlap: per-pair p<q test needed
                    }
                    if ((serial++ % nshards) != shard) continue;
                    WorkItem w;
                    w.i0 = (int32_t)i0; w.ni = (int32_t)((i1 - i0) / SEG_PAD);
                    w.j0 = (int32_t)j0; w.j1 = (int32_t)j1;
                    w.ea = ea; w.eb = eb; w.tri = tri; w.pad = 0;
                    items.push_back(w);
                }
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

// Per-warp hit queue.  In-range pairs are rare (0.3-3 % of the pairs swept), so binning them where they are
// found runs sqrt/div/atomic with one or two lanes alive and, inlined per register atom and unroll step, bloats
// the loop past the instruction cache (ncu: "no instruction" was the top stall).  Instead a hit is pushed as
// {d2 bits, slot} into a warp-private shared-memory stack with ballot-computed offsets, and the warp bins 32
// entries at a time with every lane busy.  Integer counts make the order of binning irrelevant.
static const int QCAP = 32 + 32 * 4;   // < 32 pending + one register-tile row of pushes (R <= 4)

__device__ __forceinline__ void bin_hit(uint2 e, const GridParams &g, unsigned int *__restrict__ sh,
                                        unsigned long long &ov, const SpillTarget &sp)
{
    const int b = bin_index(__uint_as_float(e.x), g);
    const int slot = (int)e.y;                    // bit 1: inter, bit 0: the J atom comes first in original order
    if (b < g.hs) {
        atomicAdd(&sh[slot * g.hs + b], 1u);
    } else {
        ++ov;
        const long long flat = (long long)((slot & 1) ? sp.slab_ba : sp.slab_ab) * g.hs + b;
        if (g.spill && flat < sp.cells) atomicAdd(&sp.counts[((slot & 2) ? sp.cells : 0) + flat], 1ull);
    }
}

template <int MODE, int R>
__device__ __forceinline__ void sweep_subtile(const float4 *__restrict__ sJ, const uint32_t *__restrict__ sO, int cnt,
                                              int jbase, const float (&xi)[R], const float (&yi)[R],
                                              const float (&zi)[R], const uint32_t (&mi)[R], const uint32_t (&oi)[R],
                                              int p0, bool tri, bool cross, const Lattice &L, const GridParams &g,
                                              uint2 *__restrict__ wq, int &qn, unsigned int *__restrict__ sh,
                                              unsigned long long &ov, const SpillTarget &sp)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 2
    for (int q = 0; q < cnt; ++q) {
        const float4 a = sJ[q];
        // all R distances first, ONE warp-uniform branch for the (rare) in-range work of the register tile
        float d2[R];
        bool hit[R];
        bool any = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            d2[r] = dist2<MODE>(xi[r], yi[r], zi[r], a.x, a.y, a.z, L);
            hit[r] = in_range(d2[r], g);
            any |= hit[r];
        }
        if (__any_sync(0xffffffffu, any)) {
            const uint32_t mj = __float_as_uint(a.w);
            const uint32_t oj = sO[q];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool h = hit[r] && (!tri || (p0 + r * SEG_PAD < jbase + q));
                const unsigned bal = __ballot_sync(0xffffffffu, h);
                if (bal) {
                    if (h) {
                        const uint32_t slot = (((mi[r] >> 8) == (mj >> 8)) ? 0u : 2u) | ((cross && (oi[r] > oj)) ? 1u : 0u);
                        wq[qn + __popc(bal & lt)] = make_uint2(__float_as_uint(d2[r]), slot);
                    }
                    qn += __popc(bal);
                }
            }
            if (qn >= 32) {
                __syncwarp();
                do {
                    qn -= 32;
                    bin_hit(wq[qn + lane], g, sh, ov, sp);
                } while (qn >= 32);
                __syncwarp();
            }
        }
    }
}

// counts layout (global, u64): [2][nEl*nEl][hs], index 0 = intra, 1 = inter.
template <int MODE, int R>
__global__ void __launch_bounds__(256)
full_hist_kernel(const float4 *__restrict__ atoms, const uint32_t *__restrict__ orig,
                 const WorkItem *__restrict__ items, int n_items, int *__restrict__ next_item, Lattice L,
                 GridParams g, int nEl, unsigned long long *__restrict__ counts,
                 unsigned long long *__restrict__ overflow)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sJ = reinterpret_cast<float4 *>(smem_raw);
    uint32_t *sO = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float4) * JS);
    uint2 *wq = reinterpret_cast<uint2 *>(smem_raw + (sizeof(float4) + sizeof(uint32_t)) * JS) + (threadIdx.x >> 5) * QCAP;
    unsigned int *sh = reinterpret_cast<unsigned int *>(smem_raw + (sizeof(float4) + sizeof(uint32_t)) * JS +
                                                        sizeof(uint2) * QCAP * (256 / 32));
    __shared__ int s_item;

    const int tid = threadIdx.x;
    const int nsh = 4 * g.hs;
    for (int c = tid; c < nsh; c += 256) sh[c] = 0u;
    unsigned long long ov = 0;
    const long long cells = (long long)nEl * nEl * g.hs;

    while (true) {
        __syncthreads();                      // previous item's flush done, s_item consumed
        if (tid == 0) s_item = atomicAdd(next_item, 1);
        __syncthreads();
        const int it = s_item;
        if (it >= n_items) break;
        const WorkItem w = items[it];

        float xi[R], yi[R], zi[R];
        uint32_t mi[R], oi[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < w.ni) {
                const int p = w.i0 + r * SEG_PAD + tid;
                const float4 a = atoms[p];
                xi[r] = a.x; yi[r] = a.y; zi[r] = a.z; mi[r] = __float_as_uint(a.w); oi[r] = orig[p];
            } else {
                xi[r] = yi[r] = zi[r] = __int_as_float(0x7FC00000);   // NaN: never in range
                mi[r] = PAD_META; oi[r] = 0xFFFFFFFFu;
            }
        }
        const bool cross = (w.ea != w.eb);
        const int p0 = w.i0 + tid;
        int qn = 0;                          // entries in this warp's hit queue (warp-uniform)
        SpillTarget sp;
        sp.counts = counts; sp.cells = cells; sp.slab_ab = w.ea * nEl + w.eb; sp.slab_ba = w.eb * nEl + w.ea;

        for (int js = w.j0; js < w.j1; js += JS) {
            const int cnt = min(JS, w.j1 - js);
            __syncthreads();                  // previous sub-tile fully consumed
            for (int q = tid; q < cnt; q += 256) { sJ[q] = atoms[js + q]; sO[q] = orig[js + q]; }
            __syncthreads();
            const bool tri = w.tri && js < w.i0 + w.ni * SEG_PAD;
            sweep_subtile<MODE, R>(sJ, sO, cnt, js, xi, yi, zi, mi, oi, p0, tri, cross, L, g, wq, qn, sh, ov, sp);
        }
        // bin what is left in this warp's queue (< 32 entries) before the item's counters are flushed
        __syncwarp();
        if ((tid & 31) < qn) bin_hit(wq[tid & 31], g, sh, ov, sp);
        qn = 0;
        __syncthreads();
        // flush the CTA-private counters of this item into the ordered global histogram
        const int slab_ab = w.ea * nEl + w.eb, slab_ba = w.eb * nEl + w.ea;
        for (int c = tid; c < nsh; c += 256) {
            const unsigned int v = sh[c];
            if (v) {
                sh[c] = 0u;
                const int slot = c / g.hs, b = c - slot * g.hs;
                const long long at = ((slot >> 1) ? cells : 0) + (long long)((slot & 1) ? slab_ba : slab_ab) * g.hs + b;
                atomicAdd(&counts[at], (unsigned long long)v);
            }
        }
    }
    if (ov) atomicAdd(overflow, ov);
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

size_t full_hist_smem_bytes(int hs)
{
    return (sizeof(float4) + sizeof(uint32_t)) * JS + sizeof(uint2) * QCAP * (256 / 32) + sizeof(unsigned int) * 4 * (size_t)hs;
}

template <int MODE, int R>
static int launch_full_t(cudaStream_t stream, int sm_count, const float4 *atoms, const uint32_t *orig,
                         const WorkItem *items, int n_items, int *next_item, const Lattice &L, const GridParams &g,
                         int nEl, unsigned long long *counts, unsigned long long *overflow)
{
    size_t smem = full_hist_smem_bytes(g.hs);
    FRMC_REQUIRE(smem <= 200 * 1024, FRMC_ELIMIT, "histSize %d needs %zu B of shared memory per CTA (limit 200 KiB)", g.hs, smem);
    auto kern = full_hist_kernel<MODE, R>;
    FRMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FRMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = std::min(n_items, sm_count * per_sm);
    if (grid < 1) return FRMC_OK;
    kern<<<grid, 256, smem, stream>>>(atoms, orig, items, n_items, next_item, L, g, nEl, counts, overflow);
    FRMC_LAUNCH_CHECK();
    return FRMC_OK;
}

// Launch the tiled kernel on prepared device arrays.  next_item must be zeroed by the caller
// (stream-ordered) before every launch.
int full_hist_launch(cudaStream_t stream, int sm_count, int mode, int R, const float4 *atoms, const uint32_t *orig,
                     const WorkItem *items, int n_items, int *next_item, const Lattice &L, const GridParams &g,
                     int nEl, unsigned long long *counts, unsigned long long *overflow)
{
#define FH_CASE(M)                                                                                         \
    case M:                                                                                                \
        return (R == 4) ? launch_full_t<M, 4>(stream, sm_count, atoms, orig, items, n_items, next_item, L, g, nEl, counts, overflow) \
                        : launch_full_t<M, 1>(stream, sm_count, atoms, orig, items, n_items, next_item, L, g, nEl, counts, overflow);
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

// Tile-shape heuristic: R register atoms per thread (I-tile = 256*R) and the J-chunk
// length, chosen so that there are enough work items to balance sm_count x occupancy CTAs.
void choose_tiling(int64_t npad, int sm_count, int &R, int64_t &chunkJ)
{
    R = (npad >= 32768) ? 4 : 1;
    const double target_items = 16.0 * sm_count * 4;
    double ti = 256.0 * R;
    double cj = (double)npad * (double)npad / (2.0 * ti * target_items);
    int64_t c = (int64_t)(cj / JS) * JS;
    if (c < JS) c = JS;
    if (c > 16384) c = 16384;
    if (npad < 8192) c = 256;   // tiny systems: finest split
    chunkJ = c;
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
extern "C" int frmc_debug_work_items(int64_t n, const int32_t *el, int nEl, int shard, int nshards, int sm_count,
                                     int64_t *n_items, int64_t *n_pairs)
{
    FRMC_REQUIRE(n >= 0 && el && n_items && n_pairs, FRMC_EINVAL, "bad arguments");
    FRMC_REQUIRE(nshards >= 1 && shard >= 0 && shard < nshards, FRMC_EINVAL, "bad shard %d of %d", shard, nshards);
    std::vector<float> coords((size_t)n * 3, 0.f);
    std::vector<int32_t> mol((size_t)n, 0);
    HostLayout lay;
    int rc = build_layout(coords.data(), n, mol.data(), el, nEl, lay);
    if (rc) return rc;
    int R; int64_t chunkJ;
    choose_tiling(lay.npad, sm_count > 0 ? sm_count : 148, R, chunkJ);
    std::vector<WorkItem> items;
    build_work_items(lay, R, chunkJ, shard, nshards, items);
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
    HostLayout lay;
    int rc = build_layout(coords, n, mol, el, nEl, lay);
    if (rc) return rc;
    Lattice L;
    for (int i = 0; i < 9; ++i) L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    GridParams g = make_grid(rmin, rmax, bin, hs);
    int mode = choose_mode_from_bounds(L.b, isPBC, lay.lo, lay.hi);
    int R; int64_t chunkJ;
    choose_tiling(lay.npad, c->sm_count, R, chunkJ);
    std::vector<WorkItem> items;
    build_work_items(lay, R, chunkJ, shard, nshards, items);

    float4 *d_atoms = (float4 *)ctx_buffer(c, 0, sizeof(float) * 4 * (size_t)lay.npad);
    uint32_t *d_orig = (uint32_t *)ctx_buffer(c, 1, sizeof(uint32_t) * (size_t)lay.npad);
    WorkItem *d_items = (WorkItem *)ctx_buffer(c, 2, sizeof(WorkItem) * items.size());
    unsigned long long *d_counts = (unsigned long long *)ctx_buffer(c, 4, sizeof(unsigned long long) * (2 * cells + 2));
    float *d_out = (float *)ctx_buffer(c, 5, sizeof(float) * 2 * cells);
    if (!d_atoms || !d_orig || !d_items || !d_counts || !d_out) return FRMC_ENOMEM;
    unsigned long long *d_ov = d_counts + 2 * cells;
    int *d_next = (int *)(d_counts + 2 * cells + 1);
    FRMC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * (2 * cells + 2), c->stream));
    if (lay.npad > 0) {
        FRMC_CUDA(cudaMemcpyAsync(d_atoms, lay.rec.data(), sizeof(float) * 4 * (size_t)lay.npad, cudaMemcpyHostToDevice, c->stream));
        FRMC_CUDA(cudaMemcpyAsync(d_orig, lay.orig.data(), sizeof(uint32_t) * (size_t)lay.npad, cudaMemcpyHostToDevice, c->stream));
    }
    if (!items.empty()) {
        FRMC_CUDA(cudaMemcpyAsync(d_items, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice, c->stream));
        rc = full_hist_launch(c->stream, c->sm_count, mode, R, d_atoms, d_orig, d_items, (int)items.size(), d_next, L, g, nEl, d_counts, d_ov);
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
