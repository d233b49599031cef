// coordnum.cu -- coordination-number counts (Extensions/atomic_coordination.pyx).
//
// The reference answers "how many atoms of list S lie in the shell [lower, upper] around core atom a" with one
// Python-level call per (atom, definition): a fancy-indexed copy of the list's coordinates, pairs_distances_to_point
// (:89-108), then a counting loop (:31-48).  all_atoms_coord_number_coords (:349-376) repeats that for every atom
// and every definition it takes part in, as core (against the definition's shell list) and as shell member
// (against its core list), and adds each count to coordNumData[definition].
//
// Here the whole call is ONE launch over a flat list of tasks (core atom, atom list, shell bounds, output slot).
// Every task is cut into chunks of 2048..32768 list entries (longer when the call is large); CTAs stride over the (task, chunk) items, so one long
// list (a per-move call: k atoms against a 10^5..10^6-atom list) fills the device as well as many short ones
// (a whole-system call).  The distance is the reference's own fp32 sequence (common.cuh dist2 with the general
// wrap, IEEE sqrt), the test is the reference's `lower <= d <= upper` on the rounded distance, counts are integers:
// the result is bit-identical to the reference for any order of evaluation.  HBM-bound in the limit (16 B
// gathered + 4 B index per list entry); at the sizes of the shipped examples it is launch + PCIe latency.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <utility>
#include <vector>

#include "common.cuh"

namespace frmc {

static const int CN_THREADS = 256;
static const int CN_CHUNK_MIN = 2048;    // list entries per work item: 8 per thread when the call is small ...
static const int CN_CHUNK_MAX = 32768;   // ... up to 128 per thread when there is work for every SM anyway (the head of
                                         // an item -- task search, two barriers -- costs about as much as 8 entries)

struct CnTask {
    int32_t core;      // atom whose position is the point
    int32_t list;      // which atom list to sweep
    int32_t out;       // slot of counts[] this task adds to
    int32_t row;       // *_totdists form: row of the distance matrix; coordinates form: 1 = bounds are not finite, test the distance itself
    float lower, upper;
    float t2lo, t2hi;  // lower <= d <= upper  <=>  t2lo <= d2 < t2hi for finite bounds (IEEE sqrt is monotone; common.cuh GridParams)
};

// first task whose item range contains `item` (item_off is non-decreasing, item_off[ntasks] = n_items; tasks over
// empty lists own no items)
__device__ __forceinline__ int cn_find_task(const long long *__restrict__ item_off, int ntasks, long long item)
{
    int lo = 0, hi = ntasks;   // invariant: item_off[lo] <= item < item_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (item_off[mid] <= item) lo = mid; else hi = mid;
    }
    return lo;
}

template <int MODE, bool DISTS>
__global__ void __launch_bounds__(CN_THREADS)
coordnum_kernel(const float4 *__restrict__ atoms, const float *__restrict__ dists, long long dstride, int chunk_len,
                const CnTask *__restrict__ tasks, int ntasks, const long long *__restrict__ item_off, long long n_items,
                const long long *__restrict__ list_off, const int32_t *__restrict__ list_idx, Lattice L,
                int *__restrict__ counts)
{
    __shared__ int s_count;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int t = cn_find_task(item_off, ntasks, item);
        const CnTask task = tasks[t];
        const long long chunk = item - item_off[t];
        const long long beg = list_off[task.list] + chunk * chunk_len;
        const long long end = min(list_off[task.list + 1], beg + chunk_len);
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        float px = 0.f, py = 0.f, pz = 0.f;
        if (!DISTS) {
            const float4 p = atoms[task.core];
            px = p.x; py = p.y; pz = p.z;
        }
        int mine = 0;
        for (long long e = beg + threadIdx.x; e < end; e += CN_THREADS) {
            const long long j = list_idx[e];
            float d;
            if (DISTS) {
                d = dists[(long long)task.row * dstride + j];
            } else {
                // pairs_distances.pyx:326-341 / :414-434: point - coords[j], wrap, basis; the IEEE sqrt only when the
                // bounds have no exact d^2 equivalent
                const float4 q = atoms[j];                         // one 16-byte gather per list entry
                const float d2 = dist2<MODE>(px, py, pz, q.x, q.y, q.z, L);
                if (!task.row) {
                    mine += (d2 >= task.t2lo && d2 < task.t2hi) ? 1 : 0;
                    continue;
                }
                d = __fsqrt_rn(d2);
            }
            // atomic_coordination.pyx:44: lowerShell <= distances[i] <= upperShell (both ends inclusive; NaN fails)
            mine += (task.lower <= d && d <= task.upper) ? 1 : 0;
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_count, mine);
        __syncthreads();
        if (threadIdx.x == 0 && s_count) atomicAdd(&counts[task.out], s_count);
        __syncthreads();
    }
}

}  // namespace frmc

using namespace frmc;

extern "C" int frmc_coordination_counts(int dev, const float *coords, int64_t n, const float *basis, int isPBC,
                                        const float *distances, int64_t nrows, int64_t ntasks, const int32_t *task_core,
                                        const int32_t *task_list, const int32_t *task_out, const float *task_lower,
                                        const float *task_upper, int64_t nlists, const int64_t *list_offsets,
                                        const int32_t *list_indexes, int64_t nout, int32_t *counts)
{
    FRMC_REQUIRE(n >= 0 && n < (1ll << 31), FRMC_EINVAL, "bad atom count");
    FRMC_REQUIRE(ntasks >= 0 && ntasks < (1ll << 31) && nlists >= 0 && nlists < (1ll << 31) && nout >= 0, FRMC_EINVAL, "bad sizes");
    FRMC_REQUIRE(nout == 0 || counts, FRMC_EINVAL, "NULL counts");
    FRMC_REQUIRE(!(coords && distances), FRMC_EINVAL, "give either coordinates or a distance matrix, not both");
    const bool use_d = distances != nullptr;
    FRMC_REQUIRE(!use_d || nrows >= 1, FRMC_EINVAL, "distance matrix without rows");
    for (int64_t i = 0; i < nout; ++i) counts[i] = 0;
    if (ntasks == 0 || nout == 0) return FRMC_OK;
    FRMC_REQUIRE(coords || distances, FRMC_EINVAL, "NULL coordinates");
    FRMC_REQUIRE(task_core && task_list && task_out && task_lower && task_upper && list_offsets, FRMC_EINVAL, "NULL task arrays");
    FRMC_REQUIRE(list_offsets[0] == 0, FRMC_EINVAL, "list_offsets[0] must be 0");
    for (int64_t l = 0; l < nlists; ++l)
        FRMC_REQUIRE(list_offsets[l + 1] >= list_offsets[l], FRMC_EINVAL, "list_offsets must not decrease");
    const int64_t total = nlists ? list_offsets[nlists] : 0;
    FRMC_REQUIRE(total == 0 || list_indexes, FRMC_EINVAL, "NULL list_indexes");
    // the reference would raise IndexError on an index outside the arrays (numpy fancy indexing); negative
    // indexes are accepted there (they wrap) and normalised here
    std::vector<int32_t> idx((size_t)total);
    for (int64_t e = 0; e < total; ++e) {
        int64_t j = list_indexes[e];
        if (j < 0) j += n;
        FRMC_REQUIRE(j >= 0 && j < n, FRMC_EINVAL, "list index %d out of bounds for %lld atoms", list_indexes[e], (long long)n);
        idx[(size_t)e] = (int32_t)j;
    }
    // items of 2048 entries for small calls (a per-move call must still spread over the SMs), longer ones when the
    // call has more than ~16 items per SM at that size
    int sm_count = 148;
    {
        DeviceCtx *c0 = get_ctx(dev);
        if (!c0) return FRMC_ECUDA;
        sm_count = c0->sm_count > 0 ? c0->sm_count : 148;
    }
    long long work = 0;
    for (int64_t t = 0; t < ntasks; ++t) {
        FRMC_REQUIRE(task_list[t] >= 0 && task_list[t] < nlists, FRMC_EINVAL, "task %lld: list %d outside 0..%lld", (long long)t, task_list[t], (long long)nlists - 1);
        work += list_offsets[task_list[t] + 1] - list_offsets[task_list[t]];
    }
    long long want = work / ((long long)sm_count * 16);
    want = (want + CN_THREADS - 1) / CN_THREADS * CN_THREADS;
    const int chunk_len = (int)std::min<long long>(CN_CHUNK_MAX, std::max<long long>(CN_CHUNK_MIN, want));
    std::vector<CnTask> tasks((size_t)ntasks);
    std::vector<long long> item_off((size_t)ntasks + 1);
    long long n_items = 0;
    std::unordered_map<uint64_t, std::pair<float, float>> thresholds;
    uint64_t last_key = 0;
    std::pair<float, float> last_val(0.0f, 0.0f);
    bool have_last = false;
    for (int64_t t = 0; t < ntasks; ++t) {
        CnTask &k = tasks[(size_t)t];
        int64_t a = task_core[t];
        if (use_d) {
            FRMC_REQUIRE(a >= 0 && a < nrows, FRMC_EINVAL, "task %lld: distance row %d outside 0..%lld", (long long)t, task_core[t], (long long)nrows - 1);
        } else {
            if (a < 0) a += n;
            FRMC_REQUIRE(a >= 0 && a < n, FRMC_EINVAL, "task %lld: core atom %d out of bounds", (long long)t, task_core[t]);
        }
        FRMC_REQUIRE(task_list[t] >= 0 && task_list[t] < nlists, FRMC_EINVAL, "task %lld: list %d outside 0..%lld", (long long)t, task_list[t], (long long)nlists - 1);
        FRMC_REQUIRE(task_out[t] >= 0 && task_out[t] < nout, FRMC_EINVAL, "task %lld: output slot %d outside 0..%lld", (long long)t, task_out[t], (long long)nout - 1);
        k.core = use_d ? 0 : (int32_t)a;
        k.row = use_d ? (int32_t)a : 0;
        k.list = task_list[t];
        k.out = task_out[t];
        k.lower = task_lower[t];
        k.upper = task_upper[t];
        k.t2lo = k.t2hi = 0.0f;
        if (!use_d) {
            // d <= upper  <=>  not (d >= the float after upper): both ends become sqrt_threshold() values; NaN or
            // infinite bounds keep the comparison on the rounded distance
            if (std::isfinite(k.lower) && std::isfinite(k.upper)) {
                // a call has a handful of distinct shells and 10^5 tasks: search the thresholds once per shell
                uint32_t bl, bu;
                memcpy(&bl, &k.lower, 4); memcpy(&bu, &k.upper, 4);
                const uint64_t key = ((uint64_t)bl << 32) | bu;
                if (!have_last || key != last_key) {
                    auto it = thresholds.find(key);
                    if (it == thresholds.end())
                        it = thresholds.emplace(key, std::make_pair(sqrt_threshold(k.lower), sqrt_threshold(nextafterf(k.upper, INFINITY)))).first;
                    last_key = key; last_val = it->second; have_last = true;
                }
                k.t2lo = last_val.first;
                k.t2hi = last_val.second;
            } else {
                k.row = 1;
            }
        }
        item_off[(size_t)t] = n_items;
        const long long len = list_offsets[k.list + 1] - list_offsets[k.list];
        n_items += (len + chunk_len - 1) / chunk_len;
    }
    item_off[(size_t)ntasks] = n_items;
    if (n_items == 0) return FRMC_OK;

    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    Lattice L;
    for (int i = 0; i < 9; ++i) L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    const size_t in_bytes = use_d ? sizeof(float) * (size_t)nrows * (size_t)n : sizeof(float4) * (size_t)n;
    std::vector<float4> packed;
    if (!use_d) {
        packed.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) packed[(size_t)i] = make_float4(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], 0.0f);
    }
    float *d_in = (float *)ctx_buffer(c, 0, std::max<size_t>(in_bytes, 16));
    CnTask *d_tasks = (CnTask *)ctx_buffer(c, 1, sizeof(CnTask) * tasks.size());
    long long *d_item_off = (long long *)ctx_buffer(c, 2, sizeof(long long) * item_off.size());
    long long *d_list_off = (long long *)ctx_buffer(c, 3, sizeof(long long) * ((size_t)nlists + 1));
    int32_t *d_idx = (int32_t *)ctx_buffer(c, 4, std::max<size_t>(sizeof(int32_t) * idx.size(), 16));
    int *d_counts = (int *)ctx_buffer(c, 5, sizeof(int) * (size_t)nout);
    if (!d_in || !d_tasks || !d_item_off || !d_list_off || !d_idx || !d_counts) return FRMC_ENOMEM;
    FRMC_CUDA(cudaMemcpyAsync(d_in, use_d ? (const void *)distances : (const void *)packed.data(), in_bytes, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_tasks, tasks.data(), sizeof(CnTask) * tasks.size(), cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_item_off, item_off.data(), sizeof(long long) * item_off.size(), cudaMemcpyHostToDevice, c->stream));
    static_assert(sizeof(long long) == sizeof(int64_t), "offsets are copied as they are");
    FRMC_CUDA(cudaMemcpyAsync(d_list_off, list_offsets, sizeof(long long) * ((size_t)nlists + 1), cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_idx, idx.data(), sizeof(int32_t) * idx.size(), cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * (size_t)nout, c->stream));
    const int grid = (int)std::min<long long>(n_items, (long long)c->sm_count * 8);
    {
    CtxTimer timer(c);                               // frmc_ctx_kernel_ms: the counting kernel
#define LAUNCH_CN(M, D) coordnum_kernel<M, D><<<grid, CN_THREADS, 0, c->stream>>>(use_d ? nullptr : (const float4 *)d_in, use_d ? d_in : nullptr, (long long)n, chunk_len, \
                                  d_tasks, (int)ntasks, d_item_off, n_items, d_list_off, d_idx, L, d_counts)
    if (use_d) {
        LAUNCH_CN(MODE_IBC, true);
    } else {
        // the wrap and basis variants of common.cuh give the same bits as the general form where they apply
        switch (choose_mode(L.b, isPBC, coords, n)) {
            case MODE_IBC: LAUNCH_CN(MODE_IBC, false); break;
            case MODE_ORTHO_FAST: LAUNCH_CN(MODE_ORTHO_FAST, false); break;
            case MODE_TRI_FAST: LAUNCH_CN(MODE_TRI_FAST, false); break;
            case MODE_ORTHO_GEN: LAUNCH_CN(MODE_ORTHO_GEN, false); break;
            default: LAUNCH_CN(MODE_TRI_GEN, false); break;
        }
    }
#undef LAUNCH_CN
    }
    FRMC_LAUNCH_CHECK();
    FRMC_CUDA(cudaMemcpyAsync(counts, d_counts, sizeof(int) * (size_t)nout, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    return FRMC_OK;
}
