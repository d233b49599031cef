// storedist.cu -- the distance-constraint pre-filter on the DEVICE STORE ("can share the
// per-move pass over the device store"; Engine.py:3281-3290 runs the rigid constraints before the experimental ones on
// every step).
//
// InterMolecularDistanceConstraint.compute_before_move / compute_after_move (Constraints/DistanceConstraints.py:606-737)
// evaluate, for the k atoms of a move,  M = multiple_atomic_distances_coords(indexes, all N atoms)  and
// F = full_atomic_distances_coords(the group alone)  on the coordinates before and after the move.  The stateless
// drop-ins (atomdist.cu) receive the whole coordinate array with every call: 0.9 ms of uploads and packing around a
// 10 us kernel.  Here the atoms are the store's own records (the ones the histogram constraints move), the types and
// windows are registered once, and ONE call evaluates all four quantities:
//
//   sd_sweep_kernel   one pass over the store: every record against the k listed atoms at their stored positions
//                     (M before) and at their moved positions (M after; a record that is itself a group member is
//                     taken at ITS moved position there) -- the exact fp32 distance arithmetic and d^2 window
//                     thresholds of atomdist.cu; counts by atomics, hits (cell, listed atom, other atom, value)
//                     appended to a list
//   sd_finish_body    the LAST CTA of the sweep (ticket; one launch per move since the end of round 2): the pairs inside the group (F before / after: upper triangle in list order), then a
//                     bitonic sort of the whole hit list by (output cell, listed atom, other atom) and one thread per
//                     cell adding its run in that order -- the reference's float32 `+=` order, so the sums are
//                     bit-identical to it -- and the list is re-armed for the next call
//
// Hits are the exception by construction (the constraint keeps atoms apart); a move with more than SD_MAX_HITS of
// them fails loudly.
#include "common.cuh"
#include "layout.h"
#include "store_view.h"

#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace frmc {

enum : int { SD_INTER = 1, SD_INTRA = 2, SD_WITHIN = 4, SD_TO_UPPER = 8, SD_TO_LOWER = 16, SD_REDUCE = 32 };
static const int SD_MAX_TYPES = 16;
static const int SD_MAX_HITS = 2048;          // 24 KB of shared memory in the finishing CTA

struct SdLimits {                    // per [type_i * nT + type_a]: window and its exact d^2 thresholds
    float lower[SD_MAX_TYPES * SD_MAX_TYPES], upper[SD_MAX_TYPES * SD_MAX_TYPES];
    float t2lower[SD_MAX_TYPES * SD_MAX_TYPES], t2upper[SD_MAX_TYPES * SD_MAX_TYPES];
};

struct SdMove {                      // by value
    int k;
    int pos[FRMC_MAX_GROUP];
    float moved[3 * FRMC_MAX_GROUP];
};

struct SdDev {                       // device state of one registered constraint
    int nT, flags;
    const int *type_pos;             // type of the record at every store position
    const SdLimits *lim;
    int *counts;                     // [4][2][nT*nT]: M before, F before, M after, F after; intra then inter
    float *sums;
    unsigned int *n_hits;            // [0] hits appended, [1] overflow flag, [2] ticket of the sweep's CTAs
    unsigned long long *keys;        // [SD_MAX_HITS] (cell << 40) | (listed atom << 32) | other atom
    float *vals;
    int *out;                        // mapped pinned host memory: [8*nT*nT] counts | [8*nT*nT] sums (bits) | overflow flag | sequence
};

__device__ __forceinline__ float sd_reduce(float d, float lower, float upper, int flags)
{
    if (flags & SD_TO_UPPER) return fabsf(__fsub_rn(upper, d));
    if (flags & SD_TO_LOWER) return fabsf(__fsub_rn(lower, d));
    if (flags & SD_REDUCE) return (d > __fdiv_rn(__fadd_rn(lower, upper), 2.0f)) ? fabsf(__fsub_rn(upper, d)) : fabsf(__fsub_rn(lower, d));
    return d;
}

// one pair (listed atom a of type ta, other atom i of type ti): window test, count, hit
__device__ __forceinline__ void sd_pair(float d2, bool same, int ta, int ti, int set, int t_listed, unsigned int other, const SdDev &S,
                                        const float *s_lo, const float *s_up, const float *s_t2lo, const float *s_t2up)
{
    const int nT = S.nT, flags = S.flags;
    if (same ? !(flags & SD_INTRA) : !(flags & SD_INTER)) return;
    const int w = ti * nT + ta;                                   // limits are indexed [type_i, type_a] (atomic_distances.pyx:77-78)
    const bool in_window = (d2 >= s_t2lo[w]) && (d2 < s_t2up[w]);
    const bool is_nan = d2 != d2;
    const bool hit = (flags & SD_WITHIN) ? (in_window || is_nan) : !in_window;
    if (!hit) return;
    const float d = sd_reduce(__fsqrt_rn(d2), s_lo[w], s_up[w], flags);
    const int cells = nT * nT;
    const int cell = (set * 2 + (same ? 0 : 1)) * cells + ta * nT + ti;   // outputs are indexed [type_a, type_i] (:111-117)
    atomicAdd(&S.counts[cell], 1);
    const unsigned int at = atomicAdd(S.n_hits, 1u);
    if (at < (unsigned)SD_MAX_HITS) {
        S.keys[at] = ((unsigned long long)(unsigned)cell << 40) | ((unsigned long long)(unsigned)t_listed << 32) | (unsigned long long)other;
        S.vals[at] = d;
    } else {
        S.n_hits[1] = 1u;
    }
}

template <int MODE>
__device__ __forceinline__ void sd_finish_body(const SdMove &mv, const Lattice &L, const SdDev &S, unsigned int seq,
                                               unsigned long long *s_key, float *s_val, const float4 *sOld, const float4 *sNew,
                                               const int *sType, const float *s_lo, const float *s_up, const float *s_t2lo,
                                               const float *s_t2up);

template <int MODE>
__global__ void __launch_bounds__(256)
sd_sweep_kernel(const float4 *__restrict__ atoms, const uint32_t *__restrict__ orig, int npad, const SdMove mv, Lattice L, const SdDev S, unsigned int seq)
{
    __shared__ float4 sOld[FRMC_MAX_GROUP], sNew[FRMC_MAX_GROUP];
    __shared__ int sPos[FRMC_MAX_GROUP], sType[FRMC_MAX_GROUP];
    __shared__ float s_lo[SD_MAX_TYPES * SD_MAX_TYPES], s_up[SD_MAX_TYPES * SD_MAX_TYPES];
    __shared__ float s_t2lo[SD_MAX_TYPES * SD_MAX_TYPES], s_t2up[SD_MAX_TYPES * SD_MAX_TYPES];
    __shared__ unsigned long long s_key[SD_MAX_HITS];      // the last CTA's sort buffers
    __shared__ float s_val[SD_MAX_HITS];
    __shared__ bool s_last;
    const int k = mv.k, nT = S.nT;
    for (int t = threadIdx.x; t < nT * nT; t += blockDim.x) {
        s_lo[t] = S.lim->lower[t]; s_up[t] = S.lim->upper[t]; s_t2lo[t] = S.lim->t2lower[t]; s_t2up[t] = S.lim->t2upper[t];
    }
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const int p = mv.pos[t];
        const float4 o = atoms[p];
        sOld[t] = o;
        sNew[t] = make_float4(mv.moved[3 * t], mv.moved[3 * t + 1], mv.moved[3 * t + 2], o.w);
        sPos[t] = p; sType[t] = S.type_pos[p];
    }
    __syncthreads();
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npad; p += gridDim.x * blockDim.x) {
        const float4 a = atoms[p];
        const uint32_t ma = __float_as_uint(a.w);
        if (ma == PAD_META) continue;
        const unsigned int oa = orig[p];
        const int ti = S.type_pos[p];
        int member = -1;
        for (int t = 0; t < k; ++t) if (sPos[t] == p) member = t;
        for (int t = 0; t < k; ++t) {
            if (member == t) continue;                             // the listed atom itself
            const float4 o = sOld[t], nw = sNew[t];
            const bool same = (__float_as_uint(o.w) >> 8) == (ma >> 8);
            const int ta = sType[t];
            // before: everything at its stored position
            sd_pair(dist2<MODE>(o.x, o.y, o.z, a.x, a.y, a.z, L), same, ta, ti, 0, t, oa, S, s_lo, s_up, s_t2lo, s_t2up);
            // after: the listed atom at its moved position, and so is the other atom when it belongs to the group
            const float4 b = (member >= 0) ? sNew[member] : a;
            sd_pair(dist2<MODE>(nw.x, nw.y, nw.z, b.x, b.y, b.z, L), same, ta, ti, 2, t, oa, S, s_lo, s_up, s_t2lo, s_t2up);
        }
    }
    // the last CTA to get here finishes the call (no second launch): every other CTA's counts and hits are behind its ticket
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(S.n_hits + 2, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    sd_finish_body<MODE>(mv, L, S, seq, s_key, s_val, sOld, sNew, sType, s_lo, s_up, s_t2lo, s_t2up);
}

// the finishing step, run by the LAST CTA of the sweep (ticket): the moved atoms, their types and the windows are still in
// its shared memory
template <int MODE>
__device__ __forceinline__ void sd_finish_body(const SdMove &mv, const Lattice &L, const SdDev &S, unsigned int seq,
                                               unsigned long long *s_key, float *s_val, const float4 *sOld, const float4 *sNew,
                                               const int *sType, const float *s_lo, const float *s_up, const float *s_t2lo,
                                               const float *s_t2up)
{
    const int k = mv.k, nT = S.nT, tid = threadIdx.x;
    // F: full_atomic_distances_coords on the group alone (atomic_distances.pyx:500-567): listed atom a, others b > a in
    // list order; the "other atom" of the key is the position in the list
    for (int e = tid; e < k * k; e += blockDim.x) {
        const int a = e / k, b = e - a * k;
        if (b <= a) continue;
        const float4 oa = sOld[a], ob = sOld[b], na = sNew[a], nb = sNew[b];
        const bool same = (__float_as_uint(oa.w) >> 8) == (__float_as_uint(ob.w) >> 8);
        sd_pair(dist2<MODE>(oa.x, oa.y, oa.z, ob.x, ob.y, ob.z, L), same, sType[a], sType[b], 1, a, (unsigned)b, S, s_lo, s_up, s_t2lo, s_t2up);
        sd_pair(dist2<MODE>(na.x, na.y, na.z, nb.x, nb.y, nb.z, L), same, sType[a], sType[b], 3, a, (unsigned)b, S, s_lo, s_up, s_t2lo, s_t2up);
    }
    __threadfence();
    __syncthreads();
    const unsigned int n_raw = *reinterpret_cast<volatile unsigned int *>(S.n_hits);
    const int n = (int)min(n_raw, (unsigned)SD_MAX_HITS);
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = tid; i < m; i += blockDim.x) {
        s_key[i] = (i < n) ? __ldcg(S.keys + i) : ~0ull;
        s_val[i] = (i < n) ? __ldcg(S.vals + i) : 0.0f;
    }
    __syncthreads();
    // bitonic sort by key (keys are unique: one entry per (cell, listed atom, other atom))
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < m; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool up = (i & size) == 0;
                    const unsigned long long ki = s_key[i], kj = s_key[j];
                    if ((ki > kj) == up) {
                        s_key[i] = kj; s_key[j] = ki;
                        const float v = s_val[i]; s_val[i] = s_val[j]; s_val[j] = v;
                    }
                }
            }
            __syncthreads();
        }
    // one thread per output cell: its run of the sorted list, added in order
    const int n_cells = 8 * nT * nT;
    for (int cell = tid; cell < n_cells; cell += blockDim.x) {
        int lo = 0, hi = n;                                        // first entry with key >= cell << 40
        const unsigned long long want = (unsigned long long)(unsigned)cell << 40;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_key[mid] < want) lo = mid + 1; else hi = mid; }
        float acc = 0.0f;
        for (int e = lo; e < n && (int)(s_key[e] >> 40) == cell; ++e) acc = __fadd_rn(acc, s_val[e]);
        // results go straight to the host (mapped pinned memory); the device counters are re-armed for the next call
        S.out[cell] = __ldcg(S.counts + cell);
        S.out[n_cells + cell] = __float_as_int(acc);
        S.counts[cell] = 0;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        S.out[2 * n_cells] = (int)(*reinterpret_cast<volatile unsigned int *>(S.n_hits + 1));
        S.n_hits[0] = 0u; S.n_hits[1] = 0u; S.n_hits[2] = 0u;
        __threadfence_system();
        *reinterpret_cast<volatile int *>(S.out + 2 * n_cells + 1) = (int)seq;   // the host spins on this word
    }
}

__global__ void sd_type_pos_kernel(const int *__restrict__ type, const uint32_t *__restrict__ orig, int npad, int *__restrict__ type_pos)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npad) return;
    const uint32_t o = orig[p];
    type_pos[p] = (o == 0xFFFFFFFFu) ? 0 : type[o];
}

}  // namespace frmc

using namespace frmc;

struct SdHost {
    SdDev dev;
    std::vector<void *> owned;
    int *h_out = nullptr;            // mapped pinned: counts | sums | overflow flag | sequence
    unsigned int seq = 0;
    int *d_type = nullptr, *d_type_pos = nullptr;   // types by original index / by position (the latter follows the store's layout)
    int64_t n0 = 0, npad = 0;
    uint64_t layout_gen = 0;
};

static std::mutex g_sd_mu;
static std::map<frmc_store *, std::vector<SdHost>> g_sd;

namespace frmc {
void storedist_release(frmc_store *s)
{
    std::lock_guard<std::mutex> lock(g_sd_mu);
    auto it = g_sd.find(s);
    if (it == g_sd.end()) return;
    for (auto &h : it->second) {
        for (void *p : h.owned) cudaFree(p);
        if (h.h_out) cudaFreeHost(h.h_out);
    }
    g_sd.erase(it);
}
}  // namespace frmc

extern "C" int frmc_store_distance_add(frmc_store *s, const int32_t *type, int nT, const float *lowerLimit, const float *upperLimit, int flags)
{
    FRMC_REQUIRE(s && type && lowerLimit && upperLimit, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(nT >= 1 && nT <= SD_MAX_TYPES, FRMC_ELIMIT, "numberOfTypes %d outside 1..%d", nT, SD_MAX_TYPES);
    StoreView v;
    int rc = store_view(s, &v);
    if (rc) return rc;
    FRMC_REQUIRE(v.n == v.n0, FRMC_ESTATE, "atoms were removed from this store: register the constraint on a fresh store");
    for (int64_t i = 0; i < v.n0; ++i)
        FRMC_REQUIRE(type[i] >= 0 && type[i] < nT, FRMC_EINVAL, "typesIndex[%lld]=%d outside 0..%d", (long long)i, type[i], nT - 1);
    SdHost h;
    memset(&h.dev, 0, sizeof(h.dev));
    const int cells = nT * nT;
    SdLimits lim;
    memset(&lim, 0, sizeof(lim));
    for (int w = 0; w < cells; ++w) {
        lim.lower[w] = lowerLimit[w]; lim.upper[w] = upperLimit[w];
        lim.t2lower[w] = sqrt_threshold(lowerLimit[w]); lim.t2upper[w] = sqrt_threshold(upperLimit[w]);
    }
    auto alloc = [&](void **out, size_t bytes) -> int {
        FRMC_CUDA(cudaMalloc(out, bytes));
        h.owned.push_back(*out);
        FRMC_CUDA(cudaMemsetAsync(*out, 0, bytes, v.stream));
        return FRMC_OK;
    };
    int *d_type = nullptr, *d_type_pos = nullptr;
    SdLimits *d_lim = nullptr;
    if ((rc = alloc((void **)&d_type, sizeof(int) * (size_t)v.n0))) return rc;
    if ((rc = alloc((void **)&d_type_pos, sizeof(int) * (size_t)std::max<int64_t>(v.npad, 1)))) return rc;
    if ((rc = alloc((void **)&d_lim, sizeof(SdLimits)))) return rc;
    if ((rc = alloc((void **)&h.dev.counts, sizeof(int) * 8 * cells))) return rc;
    if ((rc = alloc((void **)&h.dev.sums, sizeof(float) * 8 * cells))) return rc;
    if ((rc = alloc((void **)&h.dev.n_hits, sizeof(unsigned int) * 4))) return rc;
    if ((rc = alloc((void **)&h.dev.keys, sizeof(unsigned long long) * SD_MAX_HITS))) return rc;
    if ((rc = alloc((void **)&h.dev.vals, sizeof(float) * SD_MAX_HITS))) return rc;
    FRMC_CUDA(cudaMemcpyAsync(d_type, type, sizeof(int) * (size_t)v.n0, cudaMemcpyHostToDevice, v.stream));
    FRMC_CUDA(cudaMemcpyAsync(d_lim, &lim, sizeof(SdLimits), cudaMemcpyHostToDevice, v.stream));
    if (v.npad > 0) {
        sd_type_pos_kernel<<<(unsigned)((v.npad + 255) / 256), 256, 0, v.stream>>>(d_type, v.orig, (int)v.npad, d_type_pos);
        FRMC_LAUNCH_CHECK();
    }
    FRMC_CUDA(cudaStreamSynchronize(v.stream));
    h.dev.nT = nT; h.dev.flags = flags; h.dev.type_pos = d_type_pos; h.dev.lim = d_lim;
    h.d_type = d_type; h.d_type_pos = d_type_pos; h.n0 = v.n0; h.npad = v.npad; h.layout_gen = v.layout_gen;
    FRMC_CUDA(cudaHostAlloc((void **)&h.h_out, sizeof(int) * (16 * cells + 2), cudaHostAllocMapped));
    memset(h.h_out, 0, sizeof(int) * (16 * cells + 2));
    h.dev.out = h.h_out;                         // unified addressing: the mapped host pointer is valid on the device
    std::lock_guard<std::mutex> lock(g_sd_mu);
    auto &list = g_sd[s];
    list.push_back(h);
    return (int)list.size() - 1;
}

extern "C" int frmc_store_distance_move(frmc_store *s, int id, const int32_t *indexes, int k, const float *moved,
                                        int32_t *counts_out, float *sums_out)
{
    FRMC_REQUIRE(s && indexes && moved && counts_out && sums_out, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(k >= 1 && k <= FRMC_MAX_GROUP, FRMC_ELIMIT, "group size %d outside 1..%d", k, FRMC_MAX_GROUP);
    int rc = store_flush(s);                     // a deferred accept / reject of the histogram constraints is applied first
    if (rc) return rc;
    StoreView v;
    if ((rc = store_view(s, &v))) return rc;
    SdHost h;
    unsigned int seq;
    {
        std::lock_guard<std::mutex> lock(g_sd_mu);
        auto it = g_sd.find(s);
        FRMC_REQUIRE(it != g_sd.end() && id >= 0 && id < (int)it->second.size(), FRMC_EINVAL, "unknown distance constraint %d", id);
        SdHost &reg = it->second[(size_t)id];
        FRMC_REQUIRE(reg.n0 == v.n0, FRMC_ESTATE, "the store was re-numbered after atoms were removed: register the constraint again");
        if (reg.layout_gen != v.layout_gen) {
            // the store was laid out again (frmc_store_set_coords): the records changed places, the types follow them
            FRMC_REQUIRE(v.npad <= reg.npad, FRMC_ESTATE, "the store grew since the constraint was registered: register it again");
            if (v.npad > 0) {
                sd_type_pos_kernel<<<(unsigned)((v.npad + 255) / 256), 256, 0, v.stream>>>(reg.d_type, v.orig, (int)v.npad, reg.d_type_pos);
                FRMC_LAUNCH_CHECK();
            }
            reg.layout_gen = v.layout_gen;
        }
        seq = ++reg.seq;
        h = reg;
    }
    SdMove mv;
    memset(&mv, 0, sizeof(mv));
    mv.k = k;
    float lo[3], hi[3];
    for (int c = 0; c < 3; ++c) { lo[c] = v.lo[c]; hi[c] = v.hi[c]; }
    for (int t = 0; t < k; ++t) {
        FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < v.n, FRMC_EINVAL, "atom index %d outside 0..%lld", indexes[t], (long long)v.n - 1);
        mv.pos[t] = v.inv[v.rel2real ? v.rel2real[indexes[t]] : indexes[t]];
        for (int c = 0; c < 3; ++c) {
            const float x = moved[3 * t + c];
            FRMC_REQUIRE(x == x && !isinf(x), FRMC_EINVAL, "moved coordinates contain NaN or Inf");
            mv.moved[3 * t + c] = x;
            lo[c] = std::min(lo[c], x); hi[c] = std::max(hi[c], x);
        }
    }
    const int mode = choose_mode_from_bounds(v.L.b, v.isPBC, lo, hi);
    const int cells = h.dev.nT * h.dev.nT;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((v.npad + 255) / 256, (int64_t)v.sm_count * 8));
#define SD_LAUNCH(M) sd_sweep_kernel<M><<<grid, 256, 0, v.stream>>>(v.atoms, v.orig, (int)v.npad, mv, v.L, h.dev, seq)
    switch (mode) {
        case MODE_IBC: SD_LAUNCH(MODE_IBC); break;
        case MODE_ORTHO_FAST: SD_LAUNCH(MODE_ORTHO_FAST); break;
        case MODE_TRI_FAST: SD_LAUNCH(MODE_TRI_FAST); break;
        case MODE_ORTHO_GEN: SD_LAUNCH(MODE_ORTHO_GEN); break;
        default: SD_LAUNCH(MODE_TRI_GEN); break;
    }
#undef SD_LAUNCH
    FRMC_LAUNCH_CHECK();
    g_launch_count += 1;
    // the finishing CTA writes the results into mapped pinned memory and then the sequence word: spin on it instead of
    // synchronising the stream (falls back to a stream query so that a failed launch surfaces)
    volatile int *seq_word = h.h_out + 16 * cells + 1;
    for (unsigned long long spins = 0; *seq_word != (int)seq; ++spins) {
        if ((spins & 0xFFFFF) == 0xFFFFF) {
            cudaError_t e = cudaStreamQuery(v.stream);
            if (e == cudaSuccess && *seq_word != (int)seq) { set_error("distance pass finished without publishing its result"); return FRMC_ECUDA; }
            if (e != cudaSuccess && e != cudaErrorNotReady) { set_error("distance pass failed: %s", cudaGetErrorString(e)); return FRMC_ECUDA; }
        }
    }
    if (h.h_out[16 * cells]) {
        set_error("more than %d pairs of one move fall in the counted range (ordered float sums are formed by one CTA)", SD_MAX_HITS);
        return FRMC_ELIMIT;
    }
    memcpy(counts_out, h.h_out, sizeof(int) * 8 * cells);
    memcpy(sums_out, h.h_out + 8 * cells, sizeof(float) * 8 * cells);
    return FRMC_OK;
}
