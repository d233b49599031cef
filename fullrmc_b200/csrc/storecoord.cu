// storecoord.cu -- the coordination-number pre-filter on the DEVICE STORE (Engine.py:3281-3290
// evaluates the rigid constraints before the experimental ones on every step).
//
// AtomicCoordinationNumberConstraint.compute_before_move / compute_after_move
// (Constraints/AtomicCoordinationConstraints.py:519-577) call multi_atoms_coord_number_coords for the k atoms of a move,
// once on the coordinates before and once on the coordinates after it (Extensions/atomic_coordination.pyx:280-313 through
// :207-240): a moved atom that is a CORE of definition d counts the atoms of d's shell list with lower_d <= distance <=
// upper_d, a moved atom that is in d's SHELL list counts d's core atoms the same way, every count is added to
// coordNumData[d] (the atom itself is not skipped: its distance 0 counts when lower_d <= 0).  The stateless drop-in
// (coordnum.cu) receives the whole coordinate array and the Python lists with every call: ~1 ms of flattening, packing and
// copies around a 13 us kernel.
//
// Here the definitions are registered once on the store whose atoms the histogram constraints move: every record carries
// two bit masks (definitions it is a core of / it is a shell member of), and ONE launch evaluates a move:
//
//   sc_sweep_kernel   one pass over the resident records; record j against moved atom a at its stored position (before)
//                     and at its moved position (after; a record that is itself a group member is taken at ITS moved
//                     position there).  For the definitions in  (core_a & shell_j) | (shell_a & core_j)  the shell test
//                     runs on d^2 against thresholds found by exact fp32 search (as in coordnum.cu); hits go to
//                     shared-memory counters, flushed once per CTA; the LAST CTA (ticket) writes the 2 x nDef counts and
//                     a sequence word into mapped pinned host memory and re-arms the counters -- no second launch, no
//                     memcpy, no stream synchronise.
//
// Counts are integers: any order gives the reference's float32 result while a cell stays below 2^24.
#include "common.cuh"
#include "layout.h"
#include "store_view.h"

#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace frmc {

static const int SC_MAX_DEFS = 32;           // one bit per definition in a record's two masks

struct ScShells {                            // per definition
    float lower[SC_MAX_DEFS], upper[SC_MAX_DEFS];
    float t2lo[SC_MAX_DEFS], t2hi[SC_MAX_DEFS];   // lower <= d <= upper  <=>  t2lo <= d2 < t2hi (finite bounds)
    unsigned int loose;                      // bit d: bounds not finite, the test runs on the rounded distance itself
};

struct ScMove {                              // by value
    int k;
    int pos[FRMC_MAX_GROUP];
    float moved[3 * FRMC_MAX_GROUP];
};

struct ScDev {
    int ndef;
    const uint2 *mask_pos;                   // {core bits, shell bits} of the record at every store position
    const ScShells *shells;
    int *counts;                             // [2][ndef]: before, after
    unsigned int *ticket;
    int *out;                                // mapped pinned host memory: [2*ndef] counts | sequence
};

__device__ __forceinline__ void sc_pair(float d2, unsigned int defs, int set, const ScShells &sh, int *s_cnt)
{
    while (defs) {
        const int d = __ffs(defs) - 1;
        defs &= defs - 1;
        bool in;
        if ((sh.loose >> d) & 1u) {
            const float r = __fsqrt_rn(d2);                       // atomic_coordination.pyx:44 on the rounded distance
            in = (sh.lower[d] <= r) && (r <= sh.upper[d]);
        } else {
            in = (d2 >= sh.t2lo[d]) && (d2 < sh.t2hi[d]);
        }
        if (in) atomicAdd(&s_cnt[set * SC_MAX_DEFS + d], 1);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256)
sc_sweep_kernel(const float4 *__restrict__ atoms, int npad, const ScMove mv, Lattice L, const ScDev S, unsigned int seq)
{
    __shared__ float4 sOld[FRMC_MAX_GROUP], sNew[FRMC_MAX_GROUP];
    __shared__ uint2 sMask[FRMC_MAX_GROUP];
    __shared__ int sPos[FRMC_MAX_GROUP];
    __shared__ ScShells sh;
    __shared__ int s_cnt[2 * SC_MAX_DEFS];
    __shared__ unsigned int s_all;
    __shared__ bool s_last;
    const int k = mv.k, tid = threadIdx.x;
    if (tid < (int)(sizeof(ScShells) / 4)) reinterpret_cast<unsigned int *>(&sh)[tid] = reinterpret_cast<const unsigned int *>(S.shells)[tid];
    if (tid < 2 * SC_MAX_DEFS) s_cnt[tid] = 0;
    if (tid == 0) s_all = 0u;
    __syncthreads();
    for (int t = tid; t < k; t += blockDim.x) {
        const int p = mv.pos[t];
        const float4 o = atoms[p];
        sOld[t] = o;
        sNew[t] = make_float4(mv.moved[3 * t], mv.moved[3 * t + 1], mv.moved[3 * t + 2], o.w);
        sPos[t] = p;
        const uint2 m = S.mask_pos[p];
        sMask[t] = m;
        atomicOr(&s_all, m.x | m.y);
    }
    __syncthreads();
    const unsigned int any_core = s_all;                             // a group outside every definition: nothing to count
    if (any_core) {
        for (int p = blockIdx.x * blockDim.x + tid; p < npad; p += gridDim.x * blockDim.x) {
            const uint2 mj = S.mask_pos[p];
            if (!(mj.x | mj.y)) continue;
            const float4 a = atoms[p];
            if (__float_as_uint(a.w) == PAD_META) continue;
            int member = -1;
            for (int t = 0; t < k; ++t) if (sPos[t] == p) member = t;
            const float4 b = (member >= 0) ? sNew[member] : a;
            for (int t = 0; t < k; ++t) {
                const uint2 mt = sMask[t];
                const unsigned int as_core = mt.x & mj.y, as_shell = mt.y & mj.x;
                if (!(as_core | as_shell)) continue;
                const float4 o = sOld[t], nw = sNew[t];
                const float d2b = dist2<MODE>(o.x, o.y, o.z, a.x, a.y, a.z, L);
                const float d2a = dist2<MODE>(nw.x, nw.y, nw.z, b.x, b.y, b.z, L);
                sc_pair(d2b, as_core, 0, sh, s_cnt); sc_pair(d2b, as_shell, 0, sh, s_cnt);
                sc_pair(d2a, as_core, 1, sh, s_cnt); sc_pair(d2a, as_shell, 1, sh, s_cnt);
            }
        }
    }
    __syncthreads();
    if (tid < 2 * SC_MAX_DEFS) {
        const int set = tid / SC_MAX_DEFS, d = tid % SC_MAX_DEFS;
        if (d < S.ndef && s_cnt[tid]) atomicAdd(&S.counts[set * S.ndef + d], s_cnt[tid]);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(S.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = tid; c < 2 * S.ndef; c += blockDim.x) {
        S.out[c] = __ldcg(S.counts + c);
        S.counts[c] = 0;                                             // re-armed for the next call
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        *S.ticket = 0u;
        __threadfence_system();
        *reinterpret_cast<volatile int *>(S.out + 2 * S.ndef) = (int)seq;   // the host spins on this word
    }
}

__global__ void sc_mask_pos_kernel(const uint2 *__restrict__ mask, const uint32_t *__restrict__ orig, int npad, uint2 *__restrict__ mask_pos)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npad) return;
    const uint32_t o = orig[p];
    mask_pos[p] = (o == 0xFFFFFFFFu) ? make_uint2(0u, 0u) : mask[o];
}

}  // namespace frmc

using namespace frmc;

struct ScHost {
    ScDev dev;
    std::vector<void *> owned;
    uint2 *d_mask = nullptr;         // by original index
    uint2 *d_mask_pos = nullptr;     // by position (rebuilt when the store is laid out again)
    int64_t n0 = 0, npad_cap = 0;
    uint64_t layout_gen = 0;
    int *h_out = nullptr;            // mapped pinned: counts | sequence
    unsigned int seq = 0;
};

static std::mutex g_sc_mu;
static std::map<frmc_store *, std::vector<ScHost>> g_sc;

namespace frmc {
void storecoord_release(frmc_store *s)
{
    std::lock_guard<std::mutex> lock(g_sc_mu);
    auto it = g_sc.find(s);
    if (it == g_sc.end()) return;
    for (auto &h : it->second) {
        for (void *p : h.owned) cudaFree(p);
        cudaFree(h.d_mask_pos);
        if (h.h_out) cudaFreeHost(h.h_out);
    }
    g_sc.erase(it);
}
}  // namespace frmc

static int sc_map_masks(ScHost &h, const StoreView &v)
{
    if (v.npad > h.npad_cap) {
        cudaFree(h.d_mask_pos);
        h.d_mask_pos = nullptr; h.npad_cap = 0;
        FRMC_CUDA(cudaMalloc((void **)&h.d_mask_pos, sizeof(uint2) * (size_t)v.npad));
        h.npad_cap = v.npad;
    }
    if (v.npad > 0) {
        sc_mask_pos_kernel<<<(unsigned)((v.npad + 255) / 256), 256, 0, v.stream>>>(h.d_mask, v.orig, (int)v.npad, h.d_mask_pos);
        FRMC_LAUNCH_CHECK();
    }
    h.dev.mask_pos = h.d_mask_pos;
    h.layout_gen = v.layout_gen;
    return FRMC_OK;
}

extern "C" int frmc_store_coordination_add(frmc_store *s, int ndef, const int64_t *core_offsets, const int32_t *core_indexes,
                                           const int64_t *shell_offsets, const int32_t *shell_indexes, const float *lower,
                                           const float *upper)
{
    FRMC_REQUIRE(s && core_offsets && shell_offsets && lower && upper, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(ndef >= 1 && ndef <= SC_MAX_DEFS, FRMC_ELIMIT, "%d coordination-number definitions outside 1..%d", ndef, SC_MAX_DEFS);
    StoreView v;
    int rc = store_view(s, &v);
    if (rc) return rc;
    FRMC_REQUIRE(v.n == v.n0, FRMC_ESTATE, "atoms were removed from this store: register the constraint on a fresh store");
    FRMC_REQUIRE(core_offsets[0] == 0 && shell_offsets[0] == 0, FRMC_EINVAL, "offsets must start at 0");
    std::vector<uint2> mask((size_t)std::max<int64_t>(v.n0, 1), make_uint2(0u, 0u));
    for (int d = 0; d < ndef; ++d) {
        for (int role = 0; role < 2; ++role) {
            const int64_t *off = role ? shell_offsets : core_offsets;
            const int32_t *idx = role ? shell_indexes : core_indexes;
            FRMC_REQUIRE(off[d + 1] >= off[d], FRMC_EINVAL, "offsets must not decrease");
            FRMC_REQUIRE(off[d + 1] == off[d] || idx, FRMC_EINVAL, "NULL index list");
            for (int64_t e = off[d]; e < off[d + 1]; ++e) {
                const int64_t a = idx[e];
                FRMC_REQUIRE(a >= 0 && a < v.n0, FRMC_EINVAL, "definition %d: atom index %lld outside 0..%lld", d, (long long)a, (long long)v.n0 - 1);
                unsigned int &m = role ? mask[(size_t)a].y : mask[(size_t)a].x;
                // a list naming an atom twice would count it twice in the reference; the masks cannot say that
                FRMC_REQUIRE(!((m >> d) & 1u), FRMC_EINVAL, "definition %d names atom %lld twice in its %s list", d, (long long)a, role ? "shell" : "core");
                m |= 1u << d;
            }
        }
    }
    ScShells sh;
    memset(&sh, 0, sizeof(sh));
    for (int d = 0; d < ndef; ++d) {
        sh.lower[d] = lower[d]; sh.upper[d] = upper[d];
        if (std::isfinite(lower[d]) && std::isfinite(upper[d])) {
            sh.t2lo[d] = sqrt_threshold(lower[d]);
            sh.t2hi[d] = sqrt_threshold(nextafterf(upper[d], INFINITY));
        } else {
            sh.loose |= 1u << d;
        }
    }
    ScHost h;
    memset(&h.dev, 0, sizeof(h.dev));
    auto alloc = [&](void **out, size_t bytes) -> int {
        FRMC_CUDA(cudaMalloc(out, bytes));
        h.owned.push_back(*out);
        FRMC_CUDA(cudaMemsetAsync(*out, 0, bytes, v.stream));
        return FRMC_OK;
    };
    ScShells *d_sh = nullptr;
    if ((rc = alloc((void **)&h.d_mask, sizeof(uint2) * mask.size()))) return rc;
    if ((rc = alloc((void **)&d_sh, sizeof(ScShells)))) return rc;
    if ((rc = alloc((void **)&h.dev.counts, sizeof(int) * 2 * SC_MAX_DEFS))) return rc;
    if ((rc = alloc((void **)&h.dev.ticket, sizeof(unsigned int) * 4))) return rc;
    FRMC_CUDA(cudaMemcpyAsync(h.d_mask, mask.data(), sizeof(uint2) * mask.size(), cudaMemcpyHostToDevice, v.stream));
    FRMC_CUDA(cudaMemcpyAsync(d_sh, &sh, sizeof(ScShells), cudaMemcpyHostToDevice, v.stream));
    h.dev.ndef = ndef; h.dev.shells = d_sh; h.n0 = v.n0;
    if ((rc = sc_map_masks(h, v))) return rc;
    FRMC_CUDA(cudaStreamSynchronize(v.stream));
    FRMC_CUDA(cudaHostAlloc((void **)&h.h_out, sizeof(int) * (2 * SC_MAX_DEFS + 1), cudaHostAllocMapped));
    memset(h.h_out, 0, sizeof(int) * (2 * SC_MAX_DEFS + 1));
    h.dev.out = h.h_out;                         // unified addressing: the mapped host pointer is valid on the device
    std::lock_guard<std::mutex> lock(g_sc_mu);
    auto &list = g_sc[s];
    list.push_back(h);
    return (int)list.size() - 1;
}

extern "C" int frmc_store_coordination_move(frmc_store *s, int id, const int32_t *indexes, int k, const float *moved, int32_t *counts_out)
{
    FRMC_REQUIRE(s && indexes && moved && counts_out, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(k >= 1 && k <= FRMC_MAX_GROUP, FRMC_ELIMIT, "group size %d outside 1..%d", k, FRMC_MAX_GROUP);
    int rc = store_flush(s);                     // a deferred accept / reject of the histogram constraints is applied first
    if (rc) return rc;
    StoreView v;
    if ((rc = store_view(s, &v))) return rc;
    ScHost h;
    unsigned int seq;
    {
        std::lock_guard<std::mutex> lock(g_sc_mu);
        auto it = g_sc.find(s);
        FRMC_REQUIRE(it != g_sc.end() && id >= 0 && id < (int)it->second.size(), FRMC_EINVAL, "unknown coordination constraint %d", id);
        ScHost &reg = it->second[(size_t)id];
        FRMC_REQUIRE(reg.n0 == v.n0, FRMC_ESTATE, "the store was re-numbered after atoms were removed: register the constraint again");
        if (reg.layout_gen != v.layout_gen && (rc = sc_map_masks(reg, v))) return rc;   // the store was laid out again
        seq = ++reg.seq;
        h = reg;
    }
    ScMove mv;
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
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((v.npad + 255) / 256, (int64_t)v.sm_count * 8));
#define SC_LAUNCH(M) sc_sweep_kernel<M><<<grid, 256, 0, v.stream>>>(v.atoms, (int)v.npad, mv, v.L, h.dev, seq)
    switch (mode) {
        case MODE_IBC: SC_LAUNCH(MODE_IBC); break;
        case MODE_ORTHO_FAST: SC_LAUNCH(MODE_ORTHO_FAST); break;
        case MODE_TRI_FAST: SC_LAUNCH(MODE_TRI_FAST); break;
        case MODE_ORTHO_GEN: SC_LAUNCH(MODE_ORTHO_GEN); break;
        default: SC_LAUNCH(MODE_TRI_GEN); break;
    }
#undef SC_LAUNCH
    FRMC_LAUNCH_CHECK();
    g_launch_count += 1;
    volatile int *seq_word = h.h_out + 2 * h.dev.ndef;
    for (unsigned long long spins = 0; *seq_word != (int)seq; ++spins) {
        if ((spins & 0xFFFFF) == 0xFFFFF) {
            cudaError_t e = cudaStreamQuery(v.stream);
            if (e == cudaSuccess && *seq_word != (int)seq) { set_error("coordination pass finished without publishing its result"); return FRMC_ECUDA; }
            if (e != cudaSuccess && e != cudaErrorNotReady) { set_error("coordination pass failed: %s", cudaGetErrorString(e)); return FRMC_ECUDA; }
        }
    }
    memcpy(counts_out, h.h_out, sizeof(int) * 2 * (size_t)h.dev.ndef);
    return FRMC_OK;
}
