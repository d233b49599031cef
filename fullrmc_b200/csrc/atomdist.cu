// atomdist.cu -- the distance-constraint kernels of Extensions/atomic_distances.pyx:
// multiple_atomic_distances_coords (:326-417) and full_atomic_distances_coords (:500-567), i.e. for every listed
// atom a and every other atom i >= start the pairs whose distance lies inside (or, countWithinLimits = 0,
// outside) the [lower, upper) window of their type pair contribute one count and their (optionally reduced)
// distance to [type_a, type_i] of the intra- or inter-molecular arrays.
//
// Parity contract: counts are integers; the distance SUMS are float32 `+=` in the reference's loop order (listed
// atom, then i ascending), which a parallel reduction cannot reproduce.  So the sweep (same exact fp32 distance
// arithmetic and d^2 thresholds as the histogram kernels) only EMITS the hits -- key (row, i), payload (slot,
// value) -- into a device list; the list is radix-sorted by key (rs_* kernels below) and one thread per output cell adds its
// entries in that order.  Hits are the exception by construction (the constraint exists to keep atoms apart);
// the list grows on demand and the call fails loudly beyond FRMC_ATOMDIST_MAX_HITS.
#include "common.cuh"
#include "layout.h"


#include <cstring>
#include <vector>

namespace frmc {

enum : int { AD_INTER = 1, AD_INTRA = 2, AD_WITHIN = 4, AD_TO_UPPER = 8, AD_TO_LOWER = 16, AD_REDUCE = 32 };
static const unsigned long long FRMC_ATOMDIST_MAX_HITS = 1ull << 28;

struct AdRow { float x, y, z; int mol; int type; int index; int start; int row; };
static const int AD_ROWS = 128;
static const int AD_MAX_TYPES = 16;

struct AdLimits {                    // per [type_i * nT + type_a]: window and its exact d^2 thresholds
    float lower[AD_MAX_TYPES * AD_MAX_TYPES], upper[AD_MAX_TYPES * AD_MAX_TYPES];
    float t2lower[AD_MAX_TYPES * AD_MAX_TYPES], t2upper[AD_MAX_TYPES * AD_MAX_TYPES];
};

template <int MODE>
__global__ void __launch_bounds__(256)
atomdist_hits_kernel(const float *__restrict__ coords, const int *__restrict__ mol, const int *__restrict__ type, long long n,
                     const int *__restrict__ idx, long long k, int allAtoms, Lattice L, const AdLimits *__restrict__ lim, int nT,
                     int flags, int *__restrict__ counts /* [2][nT*nT] */, unsigned long long *__restrict__ n_hits,
                     unsigned long long capacity, unsigned long long *__restrict__ keys, unsigned long long *__restrict__ vals)
{
    __shared__ AdRow rows[AD_ROWS];
    __shared__ float s_lo[AD_MAX_TYPES * AD_MAX_TYPES], s_up[AD_MAX_TYPES * AD_MAX_TYPES];
    __shared__ float s_t2lo[AD_MAX_TYPES * AD_MAX_TYPES], s_t2up[AD_MAX_TYPES * AD_MAX_TYPES];
    for (int t = threadIdx.x; t < nT * nT; t += blockDim.x) {
        s_lo[t] = lim->lower[t]; s_up[t] = lim->upper[t]; s_t2lo[t] = lim->t2lower[t]; s_t2up[t] = lim->t2upper[t];
    }
    const bool inter = flags & AD_INTER, intra = flags & AD_INTRA, within = flags & AD_WITHIN;
    for (long long r0 = (long long)blockIdx.y * AD_ROWS; r0 < k; r0 += (long long)gridDim.y * AD_ROWS) {
        const int nr = (int)min((long long)AD_ROWS, k - r0);
        __syncthreads();
        for (int t = threadIdx.x; t < nr; t += blockDim.x) {
            const int a = idx[r0 + t];
            AdRow r;
            r.x = coords[3 * (long long)a]; r.y = coords[3 * (long long)a + 1]; r.z = coords[3 * (long long)a + 2];
            r.mol = mol[a]; r.type = type[a]; r.index = a; r.start = allAtoms ? 0 : a; r.row = (int)(r0 + t);
            rows[t] = r;
        }
        __syncthreads();
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const float xi = coords[3 * i], yi = coords[3 * i + 1], zi = coords[3 * i + 2];
            const int mi = mol[i], ti = type[i];
            for (int t = 0; t < nr; ++t) {
                const AdRow &r = rows[t];
                if (i < r.start || i == r.index) continue;
                const bool same = (mi == r.mol);
                if (same ? !intra : !inter) continue;
                const float d2 = dist2<MODE>(r.x, r.y, r.z, xi, yi, zi, L);
                const int w = ti * nT + r.type;                    // limits are indexed [type_i, type_a] (:77-78)
                // distance >= lower && distance < upper  <=>  d2 >= t2lower && d2 < t2upper (exact, common.cuh)
                const bool in_window = (d2 >= s_t2lo[w]) && (d2 < s_t2up[w]);
                // the reference skips NaN distances nowhere: `d < lower` and `d >= upper` are both false for NaN, so a
                // NaN pair counts when countWithinLimits (and its sum becomes NaN); outside-mode skips it only if it is
                // in the window, which a NaN never is
                const bool is_nan = d2 != d2;
                const bool hit = within ? (in_window || is_nan) : !in_window;
                if (!hit) continue;
                float d = __fsqrt_rn(d2);
                const float lower = s_lo[w], upper = s_up[w];
                if (flags & AD_TO_UPPER) d = fabsf(__fsub_rn(upper, d));
                else if (flags & AD_TO_LOWER) d = fabsf(__fsub_rn(lower, d));
                else if (flags & AD_REDUCE) d = (d > __fdiv_rn(__fadd_rn(lower, upper), 2.0f)) ? fabsf(__fsub_rn(upper, d)) : fabsf(__fsub_rn(lower, d));
                const int cell = (same ? 0 : nT * nT) + r.type * nT + ti;   // outputs are indexed [type_a, type_i] (:111-117)
                atomicAdd(&counts[cell], 1);
                const unsigned long long at = atomicAdd(n_hits, 1ull);
                if (at < capacity) {
                    keys[at] = ((unsigned long long)(unsigned)r.row << 32) | (unsigned long long)(unsigned)i;
                    vals[at] = ((unsigned long long)(unsigned)cell << 32) | (unsigned long long)__float_as_uint(d);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ stable LSD radix sort of (key, value) pairs
// 8 bits per pass, three kernels per pass: digit counts per 4096-key block, exclusive scan of the [digit][block] table,
// stable scatter (rounds of 256 keys in order; inside a round __match_any_sync ranks the lanes of a warp that share a
// digit, per-warp counts order the warps).  Passes over digits every key shares (the bits above log2 N of either half
// of the (row, other) key) are skipped by the caller.
static const int RS_THREADS = 256, RS_ROUNDS = 16, RS_BLOCK = RS_THREADS * RS_ROUNDS;

__global__ void __launch_bounds__(RS_THREADS) rs_count_kernel(const unsigned long long *__restrict__ keys, unsigned long long n, int shift,
                                                              unsigned int *__restrict__ table, int n_blocks)
{
    __shared__ unsigned int cnt[256];
    cnt[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long base = (unsigned long long)blockIdx.x * RS_BLOCK;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const unsigned long long i = base + (unsigned long long)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&cnt[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    table[(size_t)threadIdx.x * n_blocks + blockIdx.x] = cnt[threadIdx.x];
}

// exclusive scan of the digit-major table [256][n_blocks] by one CTA
__global__ void __launch_bounds__(1024) rs_scan_kernel(unsigned int *__restrict__ table, long long len)
{
    __shared__ unsigned int part[1024];
    const int t = threadIdx.x;
    const long long per = (len + 1023) / 1024, a = min(len, t * per), b = min(len, a + per);
    unsigned int sum = 0;
    for (long long i = a; i < b; ++i) sum += table[i];
    part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned int v = (t >= o) ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned int run = part[t] - sum;
    for (long long i = a; i < b; ++i) { const unsigned int v = table[i]; table[i] = run; run += v; }
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ vals,
                                                                unsigned long long n, int shift, const unsigned int *__restrict__ table, int n_blocks,
                                                                unsigned long long *__restrict__ keys_out, unsigned long long *__restrict__ vals_out)
{
    __shared__ unsigned int run[256];                  // next output position of every digit for this block
    __shared__ unsigned int wcnt[RS_THREADS / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    run[tid] = table[(size_t)tid * n_blocks + blockIdx.x];
    const unsigned long long base = (unsigned long long)blockIdx.x * RS_BLOCK;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        for (int w = 0; w < RS_THREADS / 32; ++w) wcnt[w][tid] = 0u;
        __syncthreads();
        const unsigned long long i = base + (unsigned long long)r * RS_THREADS + tid;
        const bool live = i < n;
        unsigned long long k = 0, v = 0;
        unsigned d = 256u + (unsigned)lane;            // dead lanes match nobody
        if (live) { k = keys[i]; v = vals[i]; d = (unsigned)(k >> shift) & 255u; }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (live && rank == 0) wcnt[warp][d] = (unsigned)__popc(peers);
        __syncthreads();
        if (live) {
            unsigned before = run[d];
            for (int w = 0; w < warp; ++w) before += wcnt[w][d];
            keys_out[before + rank] = k; vals_out[before + rank] = v;
        }
        __syncthreads();
        unsigned add = 0;
        for (int w = 0; w < RS_THREADS / 32; ++w) add += wcnt[w][tid];
        run[tid] += add;
        __syncthreads();
    }
}

// one thread per output cell walks the (row, i)-sorted hits and adds its own in that order
__global__ void atomdist_sum_kernel(const unsigned long long *__restrict__ vals, unsigned long long n_hits, int cells,
                                    float *__restrict__ sums)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    float acc = 0.0f;
    for (unsigned long long e = 0; e < n_hits; ++e) {
        const unsigned long long v = vals[e];
        if ((int)(v >> 32) == cell) acc = __fadd_rn(acc, __uint_as_float((unsigned)v));
    }
    sums[cell] = acc;
}

// Culled variant for the full pass in within-limits mode: the windows are a few angstroms wide, so on the k-d
// ordered store (layout.h) almost every block pair is out of reach of the largest upper limit and is never
// listed (fullhist.cu: build_pair_lists).  One CTA per item; thread t holds I record t of the item's block and
// sweeps the staged J blocks.  A pair (p, q) is the reference's (a, i) with a = the smaller ORIGINAL index (the
// full pass lists atom a against i > a), which fixes both the limits entry [type_i, type_a] and the output cell
// [type_a, type_i]; the hit's key (a, i) puts it at the reference's place in the summation order.
template <int MODE>
__global__ void __launch_bounds__(256)
atomdist_block_kernel(const float4 *__restrict__ atoms, const uint32_t *__restrict__ orig, const WorkItem *__restrict__ rows,
                      const uint32_t *__restrict__ entries, const int4 *__restrict__ items, int n_items, Lattice L,
                      const AdLimits *__restrict__ lim, int nT, int flags, int *__restrict__ counts,
                      unsigned long long *__restrict__ n_hits, unsigned long long capacity, unsigned long long *__restrict__ keys,
                      unsigned long long *__restrict__ vals)
{
    __shared__ float4 sJ[SEG_PAD];
    __shared__ uint32_t sO[SEG_PAD];
    __shared__ float s_lo[AD_MAX_TYPES * AD_MAX_TYPES], s_up[AD_MAX_TYPES * AD_MAX_TYPES];
    __shared__ float s_t2lo[AD_MAX_TYPES * AD_MAX_TYPES], s_t2up[AD_MAX_TYPES * AD_MAX_TYPES];
    const int tid = threadIdx.x;
    for (int t = tid; t < nT * nT; t += blockDim.x) {
        s_lo[t] = lim->lower[t]; s_up[t] = lim->upper[t]; s_t2lo[t] = lim->t2lower[t]; s_t2up[t] = lim->t2upper[t];
    }
    const bool inter = flags & AD_INTER, intra = flags & AD_INTRA;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int4 item = items[it];
        const WorkItem w = rows[item.x];
        const int p = w.i0 + tid;
        const float4 a = atoms[p];
        const uint32_t ma = __float_as_uint(a.w), oa = orig[p];
        const int ta = (int)(ma & 0xFFu);
        for (int e = 0; e < item.z; ++e) {
            const int jb = (int)entries[item.y + e] * SEG_PAD;
            __syncthreads();
            sJ[tid] = atoms[jb + tid]; sO[tid] = orig[jb + tid];
            __syncthreads();
            if (ma == PAD_META) continue;
            const bool tri = (w.ea == w.eb) && jb < w.i0 + w.ni * SEG_PAD;
            for (int q = 0; q < SEG_PAD; ++q) {
                const float4 c = sJ[q];
                const uint32_t mc = __float_as_uint(c.w);
                if (mc == PAD_META) continue;
                if (tri && !(p < jb + q)) continue;
                const bool same = (ma >> 8) == (mc >> 8);
                if (same ? !intra : !inter) continue;
                const float d2 = dist2<MODE>(a.x, a.y, a.z, c.x, c.y, c.z, L);
                const uint32_t oc = sO[q];
                const int tc = (int)(mc & 0xFFu);
                const bool a_first = oa < oc;                           // the reference's listed atom is the smaller original index
                const int t_a = a_first ? ta : tc, t_i = a_first ? tc : ta;
                const int wdx = t_i * nT + t_a;
                const bool in_window = (d2 >= s_t2lo[wdx]) && (d2 < s_t2up[wdx]);
                if (!(in_window || d2 != d2)) continue;
                float d = __fsqrt_rn(d2);
                const float lower = s_lo[wdx], upper = s_up[wdx];
                if (flags & AD_TO_UPPER) d = fabsf(__fsub_rn(upper, d));
                else if (flags & AD_TO_LOWER) d = fabsf(__fsub_rn(lower, d));
                else if (flags & AD_REDUCE) d = (d > __fdiv_rn(__fadd_rn(lower, upper), 2.0f)) ? fabsf(__fsub_rn(upper, d)) : fabsf(__fsub_rn(lower, d));
                const int cell = (same ? 0 : nT * nT) + t_a * nT + t_i;
                atomicAdd(&counts[cell], 1);
                const unsigned long long at = atomicAdd(n_hits, 1ull);
                if (at < capacity) {
                    keys[at] = ((unsigned long long)(a_first ? oa : oc) << 32) | (unsigned long long)(a_first ? oc : oa);
                    vals[at] = ((unsigned long long)(unsigned)cell << 32) | (unsigned long long)__float_as_uint(d);
                }
            }
        }
    }
}

}  // namespace frmc

using namespace frmc;

// grow-only hit-list scratch per device (keys, values, their ping-pong copies, the sort's digit table)
struct AdScratch {
    unsigned long long *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr;
    unsigned int *table = nullptr;
    size_t cap = 0;
};
static AdScratch g_ad[64];

static int ad_reserve(AdScratch &sc, size_t cap, cudaStream_t stream)
{
    if (sc.cap >= cap) return FRMC_OK;
    FRMC_CUDA(cudaStreamSynchronize(stream));
    cudaFree(sc.keys); cudaFree(sc.vals); cudaFree(sc.keys2); cudaFree(sc.vals2); cudaFree(sc.table);
    sc = AdScratch();
    FRMC_CUDA(cudaMalloc((void **)&sc.keys, sizeof(unsigned long long) * cap));
    FRMC_CUDA(cudaMalloc((void **)&sc.vals, sizeof(unsigned long long) * cap));
    FRMC_CUDA(cudaMalloc((void **)&sc.keys2, sizeof(unsigned long long) * cap));
    FRMC_CUDA(cudaMalloc((void **)&sc.vals2, sizeof(unsigned long long) * cap));
    FRMC_CUDA(cudaMalloc((void **)&sc.table, sizeof(unsigned int) * 256 * ((cap + RS_BLOCK - 1) / RS_BLOCK + 1)));
    sc.cap = cap;
    return FRMC_OK;
}

// sort the hits by (listed atom, other atom), add them per output cell in that order, hand the arrays back
// sort `hits` (key, value) pairs by key; index_bits = bits either 32-bit half of a key can use.  Returns the buffer
// that holds the sorted values.
static int index_bits_of(int64_t n)
{
    int bits = 1;
    while (bits < 32 && (1ll << bits) < n) ++bits;
    return bits;
}

static int ad_sort(cudaStream_t stream, AdScratch &sc, unsigned long long hits, int index_bits, const unsigned long long **sorted_vals)
{
    unsigned long long *ka = sc.keys, *va = sc.vals, *kb = sc.keys2, *vb = sc.vals2;
    const int n_blocks = (int)((hits + RS_BLOCK - 1) / RS_BLOCK);
    const int passes_half = (index_bits + 7) / 8;
    for (int half = 0; half < 2; ++half)
        for (int p = 0; p < passes_half; ++p) {
            const int shift = 32 * half + 8 * p;
            rs_count_kernel<<<n_blocks, RS_THREADS, 0, stream>>>(ka, hits, shift, sc.table, n_blocks);
            FRMC_LAUNCH_CHECK();
            rs_scan_kernel<<<1, 1024, 0, stream>>>(sc.table, 256ll * n_blocks);
            FRMC_LAUNCH_CHECK();
            rs_scatter_kernel<<<n_blocks, RS_THREADS, 0, stream>>>(ka, va, hits, shift, sc.table, n_blocks, kb, vb);
            FRMC_LAUNCH_CHECK();
            std::swap(ka, kb); std::swap(va, vb);
        }
    *sorted_vals = va;
    return FRMC_OK;
}

static int ad_finish(DeviceCtx *c, AdScratch &sc, unsigned long long hits, int index_bits, int cells, const int *d_counts, float *d_sums,
                     int32_t *nintra, float *dintra, int32_t *ninter, float *dinter)
{
    std::vector<int> h_counts((size_t)2 * cells);
    std::vector<float> h_sums((size_t)2 * cells, 0.0f);
    FRMC_CUDA(cudaMemcpyAsync(h_counts.data(), d_counts, sizeof(int) * 2 * cells, cudaMemcpyDeviceToHost, c->stream));
    if (hits > 0) {
        const unsigned long long *sorted = nullptr;
        int rc = ad_sort(c->stream, sc, hits, index_bits, &sorted);
        if (rc) return rc;
        atomdist_sum_kernel<<<(2 * cells + 63) / 64, 64, 0, c->stream>>>(sorted, hits, 2 * cells, d_sums);
        FRMC_LAUNCH_CHECK();
        FRMC_CUDA(cudaMemcpyAsync(h_sums.data(), d_sums, sizeof(float) * 2 * cells, cudaMemcpyDeviceToHost, c->stream));
    }
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    for (int w = 0; w < cells; ++w) {
        nintra[w] = h_counts[(size_t)w]; ninter[w] = h_counts[(size_t)cells + w];
        dintra[w] = h_sums[(size_t)w]; dinter[w] = h_sums[(size_t)cells + w];
    }
    return FRMC_OK;
}

extern "C" int frmc_multiple_atomic_distances_coords(int dev, const int32_t *indexes, int64_t k, const float *coords, int64_t n,
                                                     const float *basis, int isPBC, const int32_t *mol, const int32_t *type, int nT,
                                                     const float *lowerLimit, const float *upperLimit, int flags, int allAtoms,
                                                     int32_t *nintra, float *dintra, int32_t *ninter, float *dinter)
{
    FRMC_REQUIRE(n >= 0 && k >= 0, FRMC_EINVAL, "negative size");
    FRMC_REQUIRE(nT >= 1 && nT <= AD_MAX_TYPES, FRMC_ELIMIT, "numberOfElements %d outside 1..%d", nT, AD_MAX_TYPES);
    FRMC_REQUIRE(lowerLimit && upperLimit && nintra && dintra && ninter && dinter, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(n == 0 || (coords && mol && type), FRMC_EINVAL, "NULL input array");
    FRMC_REQUIRE(k == 0 || indexes, FRMC_EINVAL, "NULL indexes");
    FRMC_REQUIRE(n < (1ll << 32) && k < (1ll << 32), FRMC_ELIMIT, "more than 2^32 atoms or rows");
    const int cells = nT * nT;
    for (int64_t i = 0; i < n; ++i)
        FRMC_REQUIRE(type[i] >= 0 && type[i] < nT, FRMC_EINVAL, "elementIndex[%lld]=%d outside 0..%d", (long long)i, type[i], nT - 1);
    for (int64_t t = 0; t < k; ++t)
        FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < n, FRMC_EINVAL, "indexes[%lld]=%d outside 0..%lld", (long long)t, indexes[t], (long long)n - 1);
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    memset(nintra, 0, sizeof(int32_t) * cells); memset(ninter, 0, sizeof(int32_t) * cells);
    memset(dintra, 0, sizeof(float) * cells); memset(dinter, 0, sizeof(float) * cells);
    if (k == 0 || n == 0) return FRMC_OK;

    Lattice L;
    for (int i = 0; i < 9; ++i) L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    const int mode = choose_mode(L.b, isPBC, coords, n);
    AdLimits lim;
    for (int w = 0; w < cells; ++w) {
        lim.lower[w] = lowerLimit[w]; lim.upper[w] = upperLimit[w];
        lim.t2lower[w] = sqrt_threshold(lowerLimit[w]); lim.t2upper[w] = sqrt_threshold(upperLimit[w]);
    }
    float *d_coords = (float *)ctx_buffer(c, 0, sizeof(float) * 3 * n);
    int *d_mol = (int *)ctx_buffer(c, 1, sizeof(int) * n);
    int *d_type = (int *)ctx_buffer(c, 2, sizeof(int) * n);
    int *d_idx = (int *)ctx_buffer(c, 3, sizeof(int) * k);
    int *d_counts = (int *)ctx_buffer(c, 4, sizeof(int) * 2 * cells + 16);
    AdLimits *d_lim = (AdLimits *)ctx_buffer(c, 5, sizeof(AdLimits));
    float *d_sums = (float *)ctx_buffer(c, 6, sizeof(float) * 2 * cells);
    if (!d_coords || !d_mol || !d_type || !d_idx || !d_counts || !d_lim || !d_sums) return FRMC_ENOMEM;
    unsigned long long *d_nhits = (unsigned long long *)(d_counts + 2 * cells + (2 * cells) % 2);
    FRMC_CUDA(cudaMemcpyAsync(d_coords, coords, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_mol, mol, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_type, type, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_idx, indexes, sizeof(int) * k, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_lim, &lim, sizeof(AdLimits), cudaMemcpyHostToDevice, c->stream));

    AdScratch &sc = g_ad[dev & 63];
    int rc = ad_reserve(sc, std::max<size_t>(sc.cap, 1u << 20), c->stream);
    if (rc) return rc;
    const long long chunks = (k + AD_ROWS - 1) / AD_ROWS;
    const long long want = (n + 255) / 256, capx = (long long)c->sm_count * 8;
    dim3 grid((unsigned)std::min(want, capx), (unsigned)std::min<long long>(chunks, 4096));
    unsigned long long hits = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        FRMC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * 2 * cells + 16, c->stream));
#define LAUNCH_AD(M) atomdist_hits_kernel<M><<<grid, 256, 0, c->stream>>>(d_coords, d_mol, d_type, n, d_idx, k, allAtoms, L, d_lim, nT, flags, \
                                                                              d_counts, d_nhits, (unsigned long long)sc.cap, sc.keys, sc.vals)
        switch (mode) {
            case MODE_IBC: LAUNCH_AD(MODE_IBC); break;
            case MODE_ORTHO_FAST: LAUNCH_AD(MODE_ORTHO_FAST); break;
            case MODE_TRI_FAST: LAUNCH_AD(MODE_TRI_FAST); break;
            case MODE_ORTHO_GEN: LAUNCH_AD(MODE_ORTHO_GEN); break;
            default: LAUNCH_AD(MODE_TRI_GEN); break;
        }
#undef LAUNCH_AD
        FRMC_LAUNCH_CHECK();
        FRMC_CUDA(cudaMemcpyAsync(&hits, d_nhits, sizeof(hits), cudaMemcpyDeviceToHost, c->stream));
        FRMC_CUDA(cudaStreamSynchronize(c->stream));
        if (hits <= sc.cap) break;
        FRMC_REQUIRE(hits <= FRMC_ATOMDIST_MAX_HITS && attempt == 0, FRMC_ELIMIT,
                     "%llu pairs fall in the counted range; the ordered float sums are limited to %llu", hits, FRMC_ATOMDIST_MAX_HITS);
        rc = ad_reserve(sc, (size_t)(hits + hits / 8), c->stream);      // second pass with room for all of them
        if (rc) return rc;
    }
    return ad_finish(c, sc, hits, index_bits_of(std::max<int64_t>(n, k)), cells, d_counts, d_sums, nintra, dintra, ninter, dinter);
}

// the full pass on the k-d ordered store with block culling (within-limits mode only: there the hits are the pairs
// closer than the largest upper limit); returns 1 when the case does not qualify and the plain rows sweep must run
static int full_culled(int dev, const float *coords, int64_t n, const float *basis, int isPBC, const int32_t *mol,
                       const int32_t *type, int nT, const float *lowerLimit, const float *upperLimit, int flags,
                       int32_t *nintra, float *dintra, int32_t *ninter, float *dinter)
{
    if (!(flags & AD_WITHIN) || g_no_cull || n < 4096) return 1;
    const int cells = nT * nT;
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    Lattice L;
    for (int i = 0; i < 9; ++i) L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    AdLimits lim;
    float t2cut = 0.0f;
    for (int w = 0; w < cells; ++w) {
        lim.lower[w] = lowerLimit[w]; lim.upper[w] = upperLimit[w];
        lim.t2lower[w] = sqrt_threshold(lowerLimit[w]); lim.t2upper[w] = sqrt_threshold(upperLimit[w]);
        t2cut = std::max(t2cut, lim.t2upper[w]);
        if (upperLimit[w] != upperLimit[w]) return 1;                    // NaN limits: let the plain sweep decide
    }
    static thread_local HostLayout lay;
    int rc = build_layout(coords, n, mol, type, nT, isPBC, lay);
    if (rc) return rc;
    if (!lay.finite) return 1;                                            // NaN coordinates count as hits in the reference: plain sweep
    const int mode = choose_mode_from_bounds(L.b, isPBC, lay.lo, lay.hi);
    GridParams g;
    memset(&g, 0, sizeof(g));
    g.t2max = t2cut;
    const CullParams cp = make_cull(L, mode, g);
    if (!cp.enabled) return 1;
    std::vector<WorkItem> rows;
    build_rows(lay, 1, 0, 1, rows);
    if (rows.empty()) return 1;

    float4 *d_atoms = (float4 *)ctx_buffer(c, 0, sizeof(float4) * (size_t)lay.npad);
    uint32_t *d_orig = (uint32_t *)ctx_buffer(c, 1, sizeof(uint32_t) * (size_t)lay.npad);
    WorkItem *d_rows = (WorkItem *)ctx_buffer(c, 2, sizeof(WorkItem) * rows.size());
    float4 *d_bbox = (float4 *)ctx_buffer(c, 3, sizeof(float4) * 18 * (size_t)(lay.npad / SEG_PAD + 1));
    int *d_counts = (int *)ctx_buffer(c, 4, sizeof(int) * 2 * cells + 16);
    AdLimits *d_lim = (AdLimits *)ctx_buffer(c, 5, sizeof(AdLimits));
    float *d_sums = (float *)ctx_buffer(c, 6, sizeof(float) * 2 * cells);
    if (!d_atoms || !d_orig || !d_rows || !d_bbox || !d_counts || !d_lim || !d_sums) return FRMC_ENOMEM;
    unsigned long long *d_nhits = (unsigned long long *)(d_counts + 2 * cells + (2 * cells) % 2);
    FRMC_CUDA(cudaMemcpyAsync(d_atoms, lay.rec.data(), sizeof(float4) * (size_t)lay.npad, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_orig, lay.orig.data(), sizeof(uint32_t) * (size_t)lay.npad, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_rows, rows.data(), sizeof(WorkItem) * rows.size(), cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_lim, &lim, sizeof(AdLimits), cudaMemcpyHostToDevice, c->stream));
    static PairLists lists[64];
    PairLists &pl = lists[dev & 63];
    rc = build_pair_lists(c->stream, d_atoms, lay.npad, d_bbox, d_rows, (int)rows.size(), cp, pl);
    if (rc) return rc;
    AdScratch &sc = g_ad[dev & 63];
    rc = ad_reserve(sc, std::max<size_t>(sc.cap, 1u << 20), c->stream);
    if (rc) return rc;
    unsigned long long hits = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        FRMC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * 2 * cells + 16, c->stream));
        if (pl.n_items > 0) {
            CtxTimer timer(c);                       // frmc_ctx_kernel_ms: the block sweep
            const int grid = std::min(pl.n_items, c->sm_count * 8);
#define LAUNCH_ADB(M) atomdist_block_kernel<M><<<grid, 256, 0, c->stream>>>(d_atoms, d_orig, d_rows, pl.entries, pl.items, pl.n_items, L, d_lim, \
                                                                            nT, flags, d_counts, d_nhits, (unsigned long long)sc.cap, sc.keys, sc.vals)
            switch (mode) {
                case MODE_IBC: LAUNCH_ADB(MODE_IBC); break;
                case MODE_ORTHO_FAST: LAUNCH_ADB(MODE_ORTHO_FAST); break;
                case MODE_TRI_FAST: LAUNCH_ADB(MODE_TRI_FAST); break;
                case MODE_ORTHO_GEN: LAUNCH_ADB(MODE_ORTHO_GEN); break;
                default: LAUNCH_ADB(MODE_TRI_GEN); break;
            }
#undef LAUNCH_ADB
            FRMC_LAUNCH_CHECK();
        }
        FRMC_CUDA(cudaMemcpyAsync(&hits, d_nhits, sizeof(hits), cudaMemcpyDeviceToHost, c->stream));
        FRMC_CUDA(cudaStreamSynchronize(c->stream));
        if (hits <= sc.cap) break;
        FRMC_REQUIRE(hits <= FRMC_ATOMDIST_MAX_HITS && attempt == 0, FRMC_ELIMIT,
                     "%llu pairs fall in the counted range; the ordered float sums are limited to %llu", hits, FRMC_ATOMDIST_MAX_HITS);
        rc = ad_reserve(sc, (size_t)(hits + hits / 8), c->stream);
        if (rc) return rc;
    }
    return ad_finish(c, sc, hits, index_bits_of(n), cells, d_counts, d_sums, nintra, dintra, ninter, dinter);
}

extern "C" int frmc_full_atomic_distances_coords(int dev, const float *coords, int64_t n, const float *basis, int isPBC,
                                                 const int32_t *mol, const int32_t *type, int nT, const float *lowerLimit,
                                                 const float *upperLimit, int flags, int32_t *nintra, float *dintra,
                                                 int32_t *ninter, float *dinter)
{
    FRMC_REQUIRE(n >= 0 && n < (1ll << 31), FRMC_EINVAL, "bad atom count");
    FRMC_REQUIRE(nT >= 1 && nT <= AD_MAX_TYPES, FRMC_ELIMIT, "numberOfElements %d outside 1..%d", nT, AD_MAX_TYPES);
    FRMC_REQUIRE(lowerLimit && upperLimit && nintra && dintra && ninter && dinter, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(n == 0 || (coords && mol && type), FRMC_EINVAL, "NULL input array");
    for (int64_t i = 0; i < n; ++i)
        FRMC_REQUIRE(type[i] >= 0 && type[i] < nT, FRMC_EINVAL, "elementIndex[%lld]=%d outside 0..%d", (long long)i, type[i], nT - 1);
    const int rc = full_culled(dev, coords, n, basis, isPBC, mol, type, nT, lowerLimit, upperLimit, flags, nintra, dintra, ninter, dinter);
    if (rc != 1) return rc;
    std::vector<int32_t> idx((size_t)n);
    for (int64_t i = 0; i < n; ++i) idx[(size_t)i] = (int32_t)i;
    return frmc_multiple_atomic_distances_coords(dev, idx.data(), n, coords, n, basis, isPBC, mol, type, nT, lowerLimit, upperLimit,
                                                 flags, 0, nintra, dintra, ninter, dinter);
}
