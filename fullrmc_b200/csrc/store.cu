// store.cu -- the stateful fast path: device-resident atom store, running integer
// histograms per r-grid, fused per-move delta pass and the constraint-level epilogue
// (G(r) / g(r) / S(Q) / chi^2) on device.
//
// Replaces compute_data / compute_before_move / compute_after_move / accept_move /
// reject_move of PairDistributionConstraint, PairCorrelationConstraint and
// StructureFactorConstraint (Constraints/PairDistributionConstraints.py:1001-1166,
// PairCorrelationConstraints.py:263-392, StructureFactorConstraints.py:933-1096).
#include "common.cuh"
#include "layout.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace frmc {

GridParams make_grid(float rmin, float rmax, float bin, int hs);
int full_hist_launch(cudaStream_t stream, int sm_count, int mode, int R, const float4 *atoms, const uint32_t *orig,
                     const WorkItem *items, int n_items, int *next_item, const Lattice &L, const GridParams &g,
                     int nEl, unsigned long long *counts, unsigned long long *overflow);
void choose_tiling(int64_t npad, int sm_count, int &R, int64_t &chunkJ);
int launch_counts64_to_float(cudaStream_t stream, const unsigned long long *counts, float *out, long long cells2);

// ------------------------------------------------------------------ device-side descriptors
struct GridDev {
    GridParams g;
    long long cells;                 // nEl*nEl*hs
    unsigned long long *counts;      // committed [2][cells]  (0 intra, 1 inter)
    int *delta;                      // staged after-minus-before [2][cells]
};

struct GridSet {
    int n;
    float t2lo, t2hi;                // union of the grids' d^2 windows (cheap first test)
    GridDev grid[FRMC_MAX_GRIDS];
};

struct Proposal {                    // device copy of the staged move
    int k;
    int pos[FRMC_MAX_GROUP];         // positions in the sorted store
    float4 oldc[FRMC_MAX_GROUP];     // current records of the group atoms
    float4 newc[FRMC_MAX_GROUP];     // same meta, moved coordinates
};

struct ProposalIn {                  // what the host sends: original indices + moved box coords
    int k;
    int idx[FRMC_MAX_GROUP];
    float moved[3 * FRMC_MAX_GROUP];
};

struct ModelDev {
    int kind, grid, n_pairs, n_out, hs, sq_exact;
    float scale;
    const int *pa, *pb;
    const float *w, *D, *sv, *pref, *shape, *expv, *wts, *gr2sq;
    float *rfun;                     // [hs]    r-space function of the staged state (G(r) or g(r))
    float *total;                    // [n_out] staged model total
};

// ------------------------------------------------------------------ kernels: proposal
__global__ void prep_proposal_kernel(const ProposalIn *__restrict__ in, const int *__restrict__ inv,
                                     const float4 *__restrict__ atoms, Proposal *__restrict__ out)
{
    int t = threadIdx.x;
    int k = in->k;
    if (t == 0) out->k = k;
    if (t < k) {
        int p = inv[in->idx[t]];
        float4 o = atoms[p];
        out->pos[t] = p;
        out->oldc[t] = o;
        out->newc[t] = make_float4(in->moved[3 * t], in->moved[3 * t + 1], in->moved[3 * t + 2], o.w);
    }
}

__device__ __forceinline__ void delta_hit(float d2, int sign, int same, int slab, const GridSet &gs, int nEl,
                                          unsigned long long &ov)
{
#pragma unroll 1
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &G = gs.grid[gi];
        if (in_range(d2, G.g)) {
            int b = bin_index(d2, G.g);
            if (b < G.g.hs) atomicAdd(&G.delta[(same ? 0 : G.cells) + (long long)slab * G.g.hs + b], sign);
            else ++ov;
        }
    }
}

// One streaming pass over the whole store: every atom record is read ONCE (coalesced
// 16-byte loads) and tested against the old and the new position of every group atom;
// the signed events (-1 old, +1 new) land in the int32 delta histograms of every grid.
// Pairs inside the group are handled by block 0 with the reference's M-F convention
// (PairDistributionConstraints.py:1053-1078: the pair (t,u), u earlier in the index
// list, survives once in slab [el_t, el_u]).
template <int MODE>
__global__ void __launch_bounds__(256)
delta_kernel(const float4 *__restrict__ atoms, int npad, const Proposal *__restrict__ prop, Lattice L,
             GridSet gs, int nEl, unsigned long long *__restrict__ overflow)
{
    __shared__ float4 sOld[FRMC_MAX_GROUP];
    __shared__ float4 sNew[FRMC_MAX_GROUP];
    __shared__ int sPos[FRMC_MAX_GROUP];
    const int k = prop->k;
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        sOld[t] = prop->oldc[t]; sNew[t] = prop->newc[t]; sPos[t] = prop->pos[t];
    }
    __syncthreads();
    unsigned long long ov = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npad; p += stride) {
        const float4 a = atoms[p];
        const uint32_t mj = __float_as_uint(a.w);
        if (mj == PAD_META) continue;
        bool ingroup = false;
        for (int t = 0; t < k; ++t) ingroup |= (sPos[t] == p);
        if (ingroup) continue;
        const int ej = mj & 0xFF;
        for (int t = 0; t < k; ++t) {
            const float4 o = sOld[t], nw = sNew[t];
            const uint32_t mt = __float_as_uint(o.w);
            const float d2o = dist2<MODE>(o.x, o.y, o.z, a.x, a.y, a.z, L);
            const float d2n = dist2<MODE>(nw.x, nw.y, nw.z, a.x, a.y, a.z, L);
            const bool ho = (d2o >= gs.t2lo) && (d2o < gs.t2hi);
            const bool hn = (d2n >= gs.t2lo) && (d2n < gs.t2hi);
            if (ho || hn) {
                const int same = (mt >> 8) == (mj >> 8);
                const int slab = (int)(mt & 0xFF) * nEl + ej;
                if (ho) delta_hit(d2o, -1, same, slab, gs, nEl, ov);
                if (hn) delta_hit(d2n, +1, same, slab, gs, nEl, ov);
            }
        }
    }
    if (blockIdx.x == 0) {
        // pairs inside the moved group: (t,u) with u < t in list order -> slab [el_t, el_u]
        for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
            const int t = e / k, u = e - t * k;
            if (u >= t) continue;
            const float4 ot = sOld[t], ou = sOld[u], nt = sNew[t], nu = sNew[u];
            const uint32_t mt = __float_as_uint(ot.w), mu = __float_as_uint(ou.w);
            const int same = (mt >> 8) == (mu >> 8);
            const int slab = (int)(mt & 0xFF) * nEl + (int)(mu & 0xFF);
            const float d2o = dist2<MODE>(ot.x, ot.y, ot.z, ou.x, ou.y, ou.z, L);
            const float d2n = dist2<MODE>(nt.x, nt.y, nt.z, nu.x, nu.y, nu.z, L);
            if ((d2o >= gs.t2lo) && (d2o < gs.t2hi)) delta_hit(d2o, -1, same, slab, gs, nEl, ov);
            if ((d2n >= gs.t2lo) && (d2n < gs.t2hi)) delta_hit(d2n, +1, same, slab, gs, nEl, ov);
        }
    }
    if (ov) atomicAdd(overflow, ov);
}

// accept: fold the staged delta into the committed counts, clear it, move the atoms
__global__ void commit_kernel(GridSet gs, float4 *__restrict__ atoms, const Proposal *__restrict__ prop)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &G = gs.grid[gi];
        for (long long c = tid; c < 2 * G.cells; c += stride) {
            const int d = G.delta[c];
            if (d) { G.counts[c] = (unsigned long long)((long long)G.counts[c] + d); G.delta[c] = 0; }
        }
    }
    if (tid < prop->k) atoms[prop->pos[tid]] = prop->newc[tid];
}

// reject: clear the staged delta
__global__ void clear_delta_kernel(GridSet gs)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &G = gs.grid[gi];
        for (long long c = tid; c < 2 * G.cells; c += stride)
            if (G.delta[c]) G.delta[c] = 0;
    }
}

// ------------------------------------------------------------------ kernels: epilogue
// r-space function of every model, one thread per (model, bin).  Mirrors the numpy
// expressions of __get_total_Gr / __get_total_gr / __get_total_Sq operation by operation
// in fp32 (numpy >= 2 scalar promotion: every scalar is fp32):
//   for pair in sorted pairs:  Gr += (wij*nij)/Dij          (PairDistributionConstraints.py:855-876)
//   Gr /= shellVolumes                                      (:878)
//   Gr  = prefactor*(Gr-1)                                  (:881)   [PDF, SQ, RSQ]
//   Gr -= shape ; Gr *= scale (when != 1)                   (:883-888) [PDF]
//   PCF: gr -= shape; if scale != 1: gr = 1 + (prefactor*(gr-1)*scale)/prefactor   (PairCorrelationConstraints.py:153-163)
__global__ void rfun_kernel(const ModelDev *__restrict__ models, int n_models, GridSet gs, int nEl)
{
    const int m = blockIdx.y;
    if (m >= n_models) return;
    const ModelDev M = models[m];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M.hs) return;
    const GridDev &G = gs.grid[M.grid];
    const unsigned long long *ci = G.counts, *ce = G.counts + G.cells;
    const int *di = G.delta, *de = G.delta + G.cells;
    float acc = 0.0f;
    for (int p = 0; p < M.n_pairs; ++p) {
        const int a = M.pa[p], b = M.pb[p];
        const long long ab = ((long long)a * nEl + b) * M.hs + r;
        float n;
        if (a == b) {
            n = __fadd_rn((float)((long long)ci[ab] + di[ab]), (float)((long long)ce[ab] + de[ab]));
        } else {
            const long long ba = ((long long)b * nEl + a) * M.hs + r;
            n = __fadd_rn((float)((long long)ci[ab] + di[ab]), (float)((long long)ci[ba] + di[ba]));
            n = __fadd_rn(n, (float)((long long)ce[ab] + de[ab]));
            n = __fadd_rn(n, (float)((long long)ce[ba] + de[ba]));
        }
        acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(M.w[p], n), M.D[p]));
    }
    acc = __fdiv_rn(acc, M.sv[r]);
    float out;
    if (M.kind == FRMC_KIND_PCF) {
        out = acc;
        if (M.shape) out = __fsub_rn(out, M.shape[r]);
        if (M.scale != 1.0f) {
            float Gr = __fmul_rn(M.pref[r], __fsub_rn(out, 1.0f));
            Gr = __fmul_rn(Gr, M.scale);
            out = __fadd_rn(1.0f, __fdiv_rn(Gr, M.pref[r]));
        }
        M.total[r] = out;
    } else {
        out = __fmul_rn(M.pref[r], __fsub_rn(acc, 1.0f));
        if (M.kind == FRMC_KIND_PDF) {
            if (M.shape) out = __fsub_rn(out, M.shape[r]);
            if (M.scale != 1.0f) out = __fmul_rn(out, M.scale);
            M.total[r] = out;
        }
    }
    M.rfun[r] = out;
}

// S(Q_m) = sum_r G(r)*M[r,m] (+1), one thread per Q, r in index order, fp32 multiply then
// fp32 add with no FMA: bit-identical to np.sum(Gr.reshape((-1,1))*Gr2SqMatrix, axis=0)
// (StructureFactorConstraints.py:772-773), which accumulates rows sequentially.
// One warp per CTA so that the nQ/32 independent chains spread over as many SMs.
__global__ void __launch_bounds__(32)
sq_kernel(const ModelDev *__restrict__ models, int n_models)
{
    const int m = blockIdx.y;
    if (m >= n_models) return;
    const ModelDev M = models[m];
    if (M.kind != FRMC_KIND_SQ && M.kind != FRMC_KIND_RSQ) return;
    const int q = blockIdx.x * 32 + threadIdx.x;
    if (blockIdx.x * 32 >= M.n_out) return;
    const int qq = min(q, M.n_out - 1);
    const float *__restrict__ col = M.gr2sq + qq;
    const float *__restrict__ G = M.rfun;
    const int hs = M.hs, nq = M.n_out;
    float acc = 0.0f;
    int r = 0;
    for (; r + 8 <= hs; r += 8) {
        float g[8], c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { g[u] = G[r + u]; c[u] = col[(long long)(r + u) * nq]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, __fmul_rn(g[u], c[u]));
    }
    for (; r < hs; ++r) acc = __fadd_rn(acc, __fmul_rn(G[r], col[(long long)r * nq]));
    float s = acc;
    if (M.kind == FRMC_KIND_SQ) {
        s = __fadd_rn(s, 1.0f);
        if (M.scale != 1.0f) s = __fadd_rn(__fmul_rn(M.scale, __fsub_rn(s, 1.0f)), 1.0f);   // scale*(Sq-1)+1  (:775-778)
    } else {
        if (M.scale != 1.0f) s = __fmul_rn(M.scale, s);                                      // (:1258-1260)
    }
    if (q < M.n_out) M.total[q] = s;
}

// numpy's pairwise float32 summation (numpy/_core/src/umath/loops_utils.h.src,
// FLOAT_pairwise_sum): blocks of <=128 summed with 8 interleaved accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), remainder added sequentially, larger arrays split
// at n/2 rounded down to a multiple of 8.  Reproduced exactly so chi^2 equals
// np.add.reduce(w*(exp-model)**2) bit for bit (PairDistributionConstraints.py:833-838).
static const int PW_MAX_LEAVES = 1024;

__device__ float pairwise_combine(const float *leafsum, int n)
{
    struct Frame { int n; int state; float left; };
    Frame st[40];
    int sp = 0, next = 0;
    st[0].n = n; st[0].state = 0; st[0].left = 0.f;
    float ret = 0.f;
    while (sp >= 0) {
        Frame &f = st[sp];
        if (f.n <= 128) { ret = leafsum[next++]; --sp; continue; }
        int n2 = f.n / 2; n2 -= n2 % 8;
        if (f.state == 0) { f.state = 1; ++sp; st[sp].n = n2; st[sp].state = 0; continue; }
        if (f.state == 1) { f.left = ret; f.state = 2; ++sp; st[sp].n = f.n - n2; st[sp].state = 0; continue; }
        ret = __fadd_rn(f.left, ret); --sp;
    }
    return ret;
}

// chi^2 of every model: one CTA (256 threads) per model.
__global__ void __launch_bounds__(256)
chi2_kernel(const ModelDev *__restrict__ models, int n_models, float *__restrict__ chi2_out)
{
    extern __shared__ float v[];             // [n_out] terms
    __shared__ int leaf_off[PW_MAX_LEAVES];
    __shared__ int leaf_len[PW_MAX_LEAVES];
    __shared__ float leaf_sum[PW_MAX_LEAVES];
    __shared__ int n_leaves;
    const int m = blockIdx.x;
    if (m >= n_models) return;
    const ModelDev M = models[m];
    const int n = M.n_out;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float d = __fsub_rn(M.expv[i], M.total[i]);
        float t = __fmul_rn(d, d);
        if (M.wts) t = __fmul_rn(M.wts[i], t);
        v[i] = t;
    }
    if (threadIdx.x == 0) {
        // enumerate the leaves of the recursion left to right
        int so[40], sn[40], sp = 0, nl = 0;
        so[0] = 0; sn[0] = n;
        while (sp >= 0) {
            int o = so[sp], c = sn[sp]; --sp;
            if (c <= 128) { leaf_off[nl] = o; leaf_len[nl] = c; ++nl; continue; }
            int n2 = c / 2; n2 -= n2 % 8;
            ++sp; so[sp] = o + n2; sn[sp] = c - n2;
            ++sp; so[sp] = o; sn[sp] = n2;
        }
        n_leaves = nl;
    }
    __syncthreads();
    const int group = threadIdx.x >> 3, lane8 = threadIdx.x & 7;
    const int nl = n_leaves;
    for (int base = 0; base < nl; base += 32) {
        const int l = base + group;
        const bool live = l < nl;
        const int off = live ? leaf_off[l] : 0, len = live ? leaf_len[l] : 0;
        // leaves shorter than 8 (only a whole array with n < 8) are summed sequentially from 0;
        // otherwise lane j owns accumulator r[j].  The shuffles sit on ONE converged code path.
        const bool big = len >= 8;
        const int main_len = big ? len - (len % 8) : 0;
        float r = 0.0f;
        if (big) {
            r = v[off + lane8];
            for (int i = 8; i < main_len; i += 8) r = __fadd_rn(r, v[off + i + lane8]);
        }
        __syncwarp();
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 4));
        float res = big ? r : 0.0f;
        if (live && lane8 == 0) {
            for (int i = main_len; i < len; ++i) res = __fadd_rn(res, v[off + i]);
            leaf_sum[l] = res;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) chi2_out[m] = pairwise_combine(leaf_sum, n);
}

}  // namespace frmc

using namespace frmc;

// ------------------------------------------------------------------ host-side store
struct ModelHost {
    ModelDev dev;                    // device pointers (owned)
    float *total_committed = nullptr;
    std::vector<void *> owned;
};

struct GridHost {
    GridDev dev;
    bool valid = false;              // committed counts hold a full histogram
};

struct frmc_store {
    DeviceCtx *ctx = nullptr;
    cudaStream_t stream = nullptr;
    int dev = 0;
    int64_t n = 0, npad = 0;
    int nEl = 0, isPBC = 0;
    Lattice L;
    float lo[3], hi[3];
    std::vector<int32_t> h_mol, h_el;
    float4 *d_atoms = nullptr;
    uint32_t *d_orig = nullptr;
    int32_t *d_inv = nullptr;
    WorkItem *d_items = nullptr;
    int n_items = 0, R = 1;
    int items_shard = -1, items_nshards = -1;   // which slice of the work list d_items holds
    int64_t chunkJ = 256;
    HostLayout lay;                  // rec freed after upload; segments + inv kept
    int *d_next = nullptr;
    unsigned long long *d_overflow = nullptr;
    std::vector<GridHost> grids;
    std::vector<ModelHost> models;
    ModelDev *d_models = nullptr;
    bool models_dirty = true;
    ProposalIn *h_prop = nullptr;    // pinned
    ProposalIn *d_prop_in = nullptr;
    Proposal *d_prop = nullptr;
    float *h_chi2 = nullptr;         // pinned, device-visible
    float prop_lo[3], prop_hi[3];
    int state = 0;                   // 0 idle, 1 proposal staged
    float chi2_staged[FRMC_MAX_MODELS];
    float chi2_committed[FRMC_MAX_MODELS];
    uint64_t overflow_total = 0;
    // optional per-kernel timing (CUDA events on the store's stream; bench.py's roofline leg)
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    struct Pending { int which; cudaEvent_t a, b; };
    std::vector<Pending> ev_pending;
    double kernel_ms[4] = {0, 0, 0, 0};
    uint64_t kernel_launches[4] = {0, 0, 0, 0};
};

enum { TIME_DELTA = 0, TIME_FULL = 1, TIME_EPILOGUE = 2, TIME_COMMIT = 3 };

static cudaEvent_t timing_begin(frmc_store *s)
{
    if (!s->timing) return nullptr;
    cudaEvent_t e;
    if (!s->ev_pool.empty()) { e = s->ev_pool.back(); s->ev_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    cudaEventRecord(e, s->stream);
    return e;
}

static void timing_end(frmc_store *s, int which, cudaEvent_t a)
{
    if (!a) return;
    cudaEvent_t b;
    if (!s->ev_pool.empty()) { b = s->ev_pool.back(); s->ev_pool.pop_back(); }
    else if (cudaEventCreate(&b) != cudaSuccess) { s->ev_pool.push_back(a); return; }
    cudaEventRecord(b, s->stream);
    s->ev_pending.push_back({which, a, b});
}

// call after the stream has been synchronised
static void timing_flush(frmc_store *s)
{
    for (auto &p : s->ev_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { s->kernel_ms[p.which] += ms; s->kernel_launches[p.which]++; }
        s->ev_pool.push_back(p.a); s->ev_pool.push_back(p.b);
    }
    s->ev_pending.clear();
}

static GridSet make_gridset(frmc_store *s)
{
    GridSet gs;
    memset(&gs, 0, sizeof(gs));
    gs.n = (int)s->grids.size();
    gs.t2lo = INFINITY; gs.t2hi = 0.f;
    for (int i = 0; i < gs.n; ++i) {
        gs.grid[i] = s->grids[i].dev;
        gs.t2lo = std::min(gs.t2lo, gs.grid[i].g.t2min);
        gs.t2hi = std::max(gs.t2hi, gs.grid[i].g.t2max);
    }
    return gs;
}

static int upload_layout(frmc_store *s, const float *coords)
{
    int rc = build_layout(coords, s->n, s->h_mol.data(), s->h_el.data(), s->nEl, s->lay);
    if (rc) return rc;
    s->npad = s->lay.npad;
    for (int c = 0; c < 3; ++c) { s->lo[c] = s->lay.lo[c]; s->hi[c] = s->lay.hi[c]; }
    if (!s->d_atoms) {
        FRMC_CUDA(cudaMalloc(&s->d_atoms, sizeof(float4) * std::max<int64_t>(s->npad, 1)));
        FRMC_CUDA(cudaMalloc(&s->d_orig, sizeof(uint32_t) * std::max<int64_t>(s->npad, 1)));
        FRMC_CUDA(cudaMalloc(&s->d_inv, sizeof(int32_t) * std::max<int64_t>(s->n, 1)));
    }
    if (s->npad > 0) {
        FRMC_CUDA(cudaMemcpyAsync(s->d_atoms, s->lay.rec.data(), sizeof(float4) * s->npad, cudaMemcpyHostToDevice, s->stream));
        FRMC_CUDA(cudaMemcpyAsync(s->d_orig, s->lay.orig.data(), sizeof(uint32_t) * s->npad, cudaMemcpyHostToDevice, s->stream));
        FRMC_CUDA(cudaMemcpyAsync(s->d_inv, s->lay.inv.data(), sizeof(int32_t) * s->n, cudaMemcpyHostToDevice, s->stream));
    }
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<float>().swap(s->lay.rec);
    std::vector<uint32_t>().swap(s->lay.orig);
    return FRMC_OK;
}

static int upload_items(frmc_store *s, int shard, int nshards)
{
    if (s->d_items && s->items_shard == shard && s->items_nshards == nshards) return FRMC_OK;
    std::vector<WorkItem> items;
    choose_tiling(s->npad, s->ctx->sm_count, s->R, s->chunkJ);
    build_work_items(s->lay, s->R, s->chunkJ, shard, nshards, items);
    if (s->d_items) { cudaFree(s->d_items); s->d_items = nullptr; }
    s->n_items = (int)items.size();
    FRMC_CUDA(cudaMalloc(&s->d_items, sizeof(WorkItem) * std::max<size_t>(items.size(), 1)));
    if (!items.empty())
        FRMC_CUDA(cudaMemcpyAsync(s->d_items, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    s->items_shard = shard; s->items_nshards = nshards;
    return FRMC_OK;
}

static int sync_models(frmc_store *s)
{
    if (!s->models_dirty) return FRMC_OK;
    std::vector<ModelDev> tmp;
    for (auto &m : s->models) tmp.push_back(m.dev);
    if (!s->d_models) FRMC_CUDA(cudaMalloc(&s->d_models, sizeof(ModelDev) * FRMC_MAX_MODELS));
    if (!tmp.empty())
        FRMC_CUDA(cudaMemcpyAsync(s->d_models, tmp.data(), sizeof(ModelDev) * tmp.size(), cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));   // tmp goes out of scope
    s->models_dirty = false;
    return FRMC_OK;
}

// totals + chi^2 of every model from (counts + delta); results land in s->h_chi2 after a stream sync
static int launch_epilogue(frmc_store *s)
{
    const int nm = (int)s->models.size();
    if (nm == 0) return FRMC_OK;
    int rc = sync_models(s);
    if (rc) return rc;
    GridSet gs = make_gridset(s);
    int max_hs = 0, max_out = 0, max_q = 0;
    for (auto &m : s->models) {
        max_hs = std::max(max_hs, m.dev.hs);
        max_out = std::max(max_out, m.dev.n_out);
        if (m.dev.kind == FRMC_KIND_SQ || m.dev.kind == FRMC_KIND_RSQ) max_q = std::max(max_q, m.dev.n_out);
    }
    cudaEvent_t t0 = timing_begin(s);
    dim3 g1((unsigned)((max_hs + 127) / 128), (unsigned)nm);
    rfun_kernel<<<g1, 128, 0, s->stream>>>(s->d_models, nm, gs, s->nEl);
    FRMC_LAUNCH_CHECK();
    if (max_q > 0) {
        dim3 g2((unsigned)((max_q + 31) / 32), (unsigned)nm);
        sq_kernel<<<g2, 32, 0, s->stream>>>(s->d_models, nm);
        FRMC_LAUNCH_CHECK();
    }
    size_t smem = sizeof(float) * (size_t)max_out;
    if (smem > 40 * 1024) FRMC_CUDA(cudaFuncSetAttribute(chi2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chi2_kernel<<<nm, 256, smem, s->stream>>>(s->d_models, nm, s->h_chi2);
    FRMC_LAUNCH_CHECK();
    timing_end(s, TIME_EPILOGUE, t0);
    return FRMC_OK;
}

template <typename T>
static int dev_copy(frmc_store *s, ModelHost &mh, const T *src, size_t count, const T **dst)
{
    *dst = nullptr;
    if (!src || count == 0) return FRMC_OK;
    void *p = nullptr;
    FRMC_CUDA(cudaMalloc(&p, sizeof(T) * count));
    mh.owned.push_back(p);
    FRMC_CUDA(cudaMemcpy(p, src, sizeof(T) * count, cudaMemcpyHostToDevice));
    *dst = (const T *)p;
    return FRMC_OK;
}

extern "C" {

frmc_store *frmc_store_create(int dev, int64_t n, const float *coords, const float *basis, int isPBC,
                              const int32_t *mol, const int32_t *el, int nEl)
{
    if (n < 1 || !coords || !mol || !el) { set_error("frmc_store_create: need n >= 1 and non-NULL arrays"); return nullptr; }
    if (isPBC && !basis) { set_error("frmc_store_create: periodic store needs a basis"); return nullptr; }
    DeviceCtx *c = get_ctx(dev);
    if (!c) return nullptr;
    frmc_store *s = new frmc_store();
    s->ctx = c; s->dev = dev; s->n = n; s->nEl = nEl; s->isPBC = isPBC ? 1 : 0;
    for (int i = 0; i < 9; ++i) s->L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    s->h_mol.assign(mol, mol + n);
    s->h_el.assign(el, el + n);
    auto fail = [&](const char *what) -> frmc_store * {
        std::string msg = std::string(what) + ": " + frmc_last_error();
        frmc_store_destroy(s);
        set_error("%s", msg.c_str());
        return nullptr;
    };
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream create");
    if (upload_layout(s, coords)) return fail("layout upload");
    if (upload_items(s, 0, 1)) return fail("work list upload");
    if (cudaMalloc(&s->d_next, sizeof(int) * 4) != cudaSuccess) return fail("alloc");
    if (cudaMalloc(&s->d_overflow, sizeof(unsigned long long)) != cudaSuccess) return fail("alloc");
    cudaMemset(s->d_overflow, 0, sizeof(unsigned long long));
    if (cudaMalloc(&s->d_prop_in, sizeof(ProposalIn)) != cudaSuccess) return fail("alloc");
    if (cudaMalloc(&s->d_prop, sizeof(Proposal)) != cudaSuccess) return fail("alloc");
    if (cudaMallocHost(&s->h_prop, sizeof(ProposalIn)) != cudaSuccess) return fail("pinned alloc");
    if (cudaHostAlloc(&s->h_chi2, sizeof(float) * FRMC_MAX_MODELS, cudaHostAllocMapped) != cudaSuccess) return fail("pinned alloc");
    for (int i = 0; i < FRMC_MAX_MODELS; ++i) { s->h_chi2[i] = 0.f; s->chi2_staged[i] = 0.f; s->chi2_committed[i] = 0.f; }
    return s;
}

void frmc_store_destroy(frmc_store *s)
{
    if (!s) return;
    cudaSetDevice(s->dev);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto &m : s->models) {
        for (void *p : m.owned) cudaFree(p);
    }
    for (auto &g : s->grids) { cudaFree(g.dev.counts); cudaFree(g.dev.delta); }
    cudaFree(s->d_atoms); cudaFree(s->d_orig); cudaFree(s->d_inv); cudaFree(s->d_items); cudaFree(s->d_next);
    cudaFree(s->d_overflow); cudaFree(s->d_models); cudaFree(s->d_prop_in); cudaFree(s->d_prop);
    for (auto &p : s->ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : s->ev_pool) cudaEventDestroy(e);
    if (s->h_prop) cudaFreeHost(s->h_prop);
    if (s->h_chi2) cudaFreeHost(s->h_chi2);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

void *frmc_store_stream(frmc_store *s) { return s ? (void *)s->stream : nullptr; }

int frmc_store_set_coords(frmc_store *s, const float *coords, const float *basis)
{
    FRMC_REQUIRE(s && coords, FRMC_EINVAL, "NULL argument");
    FRMC_CUDA(cudaSetDevice(s->dev));
    if (basis) for (int i = 0; i < 9; ++i) s->L.b[i] = basis[i];
    int rc = upload_layout(s, coords);
    if (rc) return rc;
    s->items_shard = s->items_nshards = -1;
    for (auto &g : s->grids) g.valid = false;
    s->state = 0;
    GridSet gs = make_gridset(s);
    if (gs.n) { clear_delta_kernel<<<s->ctx->sm_count, 256, 0, s->stream>>>(gs); FRMC_LAUNCH_CHECK(); }
    return FRMC_OK;
}

int frmc_store_get_coords(frmc_store *s, float *coords_out)
{
    FRMC_REQUIRE(s && coords_out, FRMC_EINVAL, "NULL argument");
    FRMC_CUDA(cudaSetDevice(s->dev));
    std::vector<float> rec((size_t)s->npad * 4);
    FRMC_CUDA(cudaMemcpyAsync(rec.data(), s->d_atoms, sizeof(float4) * s->npad, cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < s->n; ++i) {
        int64_t p = s->lay.inv[i];
        coords_out[3 * i] = rec[4 * p]; coords_out[3 * i + 1] = rec[4 * p + 1]; coords_out[3 * i + 2] = rec[4 * p + 2];
    }
    return FRMC_OK;
}

int frmc_grid_add(frmc_store *s, float rmin, float rmax, float bin, int hs)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->grids.size() < FRMC_MAX_GRIDS, FRMC_ELIMIT, "at most %d grids per store", FRMC_MAX_GRIDS);
    FRMC_REQUIRE(hs >= 1 && bin > 0.f, FRMC_EINVAL, "bad grid (hs=%d, bin=%g)", hs, bin);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "cannot add a grid while a proposal is staged");
    FRMC_CUDA(cudaSetDevice(s->dev));
    GridHost gh;
    gh.dev.g = make_grid(rmin, rmax, bin, hs);
    gh.dev.cells = (long long)s->nEl * s->nEl * hs;
    FRMC_CUDA(cudaMalloc(&gh.dev.counts, sizeof(unsigned long long) * 2 * gh.dev.cells));
    FRMC_CUDA(cudaMalloc(&gh.dev.delta, sizeof(int) * 2 * gh.dev.cells));
    FRMC_CUDA(cudaMemsetAsync(gh.dev.counts, 0, sizeof(unsigned long long) * 2 * gh.dev.cells, s->stream));
    FRMC_CUDA(cudaMemsetAsync(gh.dev.delta, 0, sizeof(int) * 2 * gh.dev.cells, s->stream));
    s->grids.push_back(gh);
    return (int)s->grids.size() - 1;
}

int frmc_model_add(frmc_store *s, int grid, const frmc_model_desc *d)
{
    FRMC_REQUIRE(s && d, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(grid >= 0 && grid < (int)s->grids.size(), FRMC_EINVAL, "unknown grid %d", grid);
    FRMC_REQUIRE(s->models.size() < FRMC_MAX_MODELS, FRMC_ELIMIT, "at most %d models per store", FRMC_MAX_MODELS);
    FRMC_REQUIRE(d->kind >= FRMC_KIND_PDF && d->kind <= FRMC_KIND_RSQ, FRMC_EINVAL, "unknown model kind %d", d->kind);
    FRMC_REQUIRE(d->n_pairs >= 1 && d->pair_a && d->pair_b && d->pair_w && d->pair_D, FRMC_EINVAL, "bad pair table");
    FRMC_REQUIRE(d->shell_volumes && d->prefactor && d->experimental && d->n_out >= 1, FRMC_EINVAL, "missing model arrays");
    const int hs = s->grids[grid].dev.g.hs;
    const bool is_sq = (d->kind == FRMC_KIND_SQ || d->kind == FRMC_KIND_RSQ);
    FRMC_REQUIRE(is_sq ? (d->gr2sq != nullptr) : (d->n_out == hs), FRMC_EINVAL,
                 "model output length %d inconsistent with grid histSize %d", d->n_out, hs);
    FRMC_REQUIRE(d->n_out <= 128 * PW_MAX_LEAVES / 2, FRMC_ELIMIT, "model output too long (%d)", d->n_out);
    for (int p = 0; p < d->n_pairs; ++p)
        FRMC_REQUIRE(d->pair_a[p] >= 0 && d->pair_a[p] < s->nEl && d->pair_b[p] >= 0 && d->pair_b[p] < s->nEl,
                     FRMC_EINVAL, "pair %d references an element outside 0..%d", p, s->nEl - 1);
    FRMC_CUDA(cudaSetDevice(s->dev));
    ModelHost mh;
    memset(&mh.dev, 0, sizeof(mh.dev));
    mh.dev.kind = d->kind; mh.dev.grid = grid; mh.dev.n_pairs = d->n_pairs; mh.dev.n_out = d->n_out;
    mh.dev.hs = hs; mh.dev.sq_exact = d->sq_exact; mh.dev.scale = d->scale;
    int rc;
    if ((rc = dev_copy(s, mh, d->pair_a, d->n_pairs, &mh.dev.pa))) return rc;
    if ((rc = dev_copy(s, mh, d->pair_b, d->n_pairs, &mh.dev.pb))) return rc;
    if ((rc = dev_copy(s, mh, d->pair_w, d->n_pairs, &mh.dev.w))) return rc;
    if ((rc = dev_copy(s, mh, d->pair_D, d->n_pairs, &mh.dev.D))) return rc;
    if ((rc = dev_copy(s, mh, d->shell_volumes, hs, &mh.dev.sv))) return rc;
    if ((rc = dev_copy(s, mh, d->prefactor, hs, &mh.dev.pref))) return rc;
    if ((rc = dev_copy(s, mh, d->shape, hs, &mh.dev.shape))) return rc;
    if ((rc = dev_copy(s, mh, d->experimental, d->n_out, &mh.dev.expv))) return rc;
    if ((rc = dev_copy(s, mh, d->data_weights, d->n_out, &mh.dev.wts))) return rc;
    if (is_sq && (rc = dev_copy(s, mh, d->gr2sq, (size_t)hs * d->n_out, &mh.dev.gr2sq))) return rc;
    void *p = nullptr;
    FRMC_CUDA(cudaMalloc(&p, sizeof(float) * hs)); mh.owned.push_back(p); mh.dev.rfun = (float *)p;
    FRMC_CUDA(cudaMalloc(&p, sizeof(float) * d->n_out)); mh.owned.push_back(p); mh.dev.total = (float *)p;
    FRMC_CUDA(cudaMalloc(&p, sizeof(float) * d->n_out)); mh.owned.push_back(p); mh.total_committed = (float *)p;
    FRMC_CUDA(cudaMemset(mh.dev.total, 0, sizeof(float) * d->n_out));
    FRMC_CUDA(cudaMemset(mh.total_committed, 0, sizeof(float) * d->n_out));
    s->models.push_back(mh);
    s->models_dirty = true;
    return (int)s->models.size() - 1;
}

int frmc_model_set_scale(frmc_store *s, int model, float scale)
{
    FRMC_REQUIRE(s && model >= 0 && model < (int)s->models.size(), FRMC_EINVAL, "unknown model %d", model);
    s->models[model].dev.scale = scale;
    s->models_dirty = true;
    return FRMC_OK;
}

static int current_mode(frmc_store *s, const float *extra_lo, const float *extra_hi)
{
    float lo[3], hi[3];
    for (int c = 0; c < 3; ++c) {
        lo[c] = extra_lo ? std::min(s->lo[c], extra_lo[c]) : s->lo[c];
        hi[c] = extra_hi ? std::max(s->hi[c], extra_hi[c]) : s->hi[c];
    }
    return choose_mode_from_bounds(s->L.b, s->isPBC, lo, hi);
}

int frmc_compute_data_shard(frmc_store *s, int shard, int nshards)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(nshards >= 1 && shard >= 0 && shard < nshards, FRMC_EINVAL, "bad shard %d of %d", shard, nshards);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_CUDA(cudaSetDevice(s->dev));
    int rc = upload_items(s, shard, nshards);
    if (rc) return rc;
    const int mode = current_mode(s, nullptr, nullptr);
    for (auto &g : s->grids) {
        FRMC_CUDA(cudaMemsetAsync(g.dev.counts, 0, sizeof(unsigned long long) * 2 * g.dev.cells, s->stream));
        FRMC_CUDA(cudaMemsetAsync(g.dev.delta, 0, sizeof(int) * 2 * g.dev.cells, s->stream));
        FRMC_CUDA(cudaMemsetAsync(s->d_next, 0, sizeof(int) * 4, s->stream));
        if (s->n_items > 0) {
            cudaEvent_t t0 = timing_begin(s);
            rc = full_hist_launch(s->stream, s->ctx->sm_count, mode, s->R, s->d_atoms, s->d_orig, s->d_items, s->n_items,
                                  s->d_next, s->L, g.dev.g, s->nEl, g.dev.counts, s->d_overflow);
            if (rc) return rc;
            timing_end(s, TIME_FULL, t0);
        }
        g.valid = true;
    }
    return FRMC_OK;
}

void *frmc_grid_counts_ptr(frmc_store *s, int grid, int64_t *n_cells)
{
    if (!s || grid < 0 || grid >= (int)s->grids.size()) { set_error("unknown grid %d", grid); return nullptr; }
    if (n_cells) *n_cells = 2 * s->grids[grid].dev.cells;
    return s->grids[grid].dev.counts;
}

int frmc_finalize_data(frmc_store *s, float *chi2)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_CUDA(cudaSetDevice(s->dev));
    int rc = launch_epilogue(s);
    if (rc) return rc;
    for (auto &m : s->models)
        FRMC_CUDA(cudaMemcpyAsync(m.total_committed, m.dev.total, sizeof(float) * m.dev.n_out, cudaMemcpyDeviceToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    timing_flush(s);
    for (size_t i = 0; i < s->models.size(); ++i) {
        s->chi2_committed[i] = s->h_chi2[i];
        if (chi2) chi2[i] = s->h_chi2[i];
    }
    return FRMC_OK;
}

int frmc_compute_data(frmc_store *s, float *chi2)
{
    int rc = frmc_compute_data_shard(s, 0, 1);
    if (rc) return rc;
    return frmc_finalize_data(s, chi2);
}

int frmc_propose(frmc_store *s, const int32_t *indexes, int k, const float *moved, float *chi2_after)
{
    FRMC_REQUIRE(s && indexes && moved, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(k >= 1 && k <= FRMC_MAX_GROUP, FRMC_ELIMIT, "group size %d outside 1..%d", k, FRMC_MAX_GROUP);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is already staged; accept or reject it first");
    FRMC_REQUIRE(!s->grids.empty(), FRMC_ESTATE, "no grid registered");
    for (auto &g : s->grids) FRMC_REQUIRE(g.valid, FRMC_ESTATE, "call frmc_compute_data before proposing moves");
    FRMC_CUDA(cudaSetDevice(s->dev));
    ProposalIn *h = s->h_prop;
    h->k = k;
    for (int c = 0; c < 3; ++c) { s->prop_lo[c] = INFINITY; s->prop_hi[c] = -INFINITY; }
    bool finite = true;
    for (int t = 0; t < k; ++t) {
        FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < s->n, FRMC_EINVAL, "atom index %d outside 0..%lld", indexes[t], (long long)s->n - 1);
        h->idx[t] = indexes[t];
        for (int c = 0; c < 3; ++c) {
            float v = moved[3 * t + c];
            h->moved[3 * t + c] = v;
            if (!(v == v) || isinf(v)) finite = false;
            s->prop_lo[c] = std::min(s->prop_lo[c], v);
            s->prop_hi[c] = std::max(s->prop_hi[c], v);
        }
    }
    FRMC_REQUIRE(finite, FRMC_EINVAL, "moved coordinates contain NaN or Inf");
    const int mode = current_mode(s, s->prop_lo, s->prop_hi);
    FRMC_CUDA(cudaMemcpyAsync(s->d_prop_in, h, sizeof(int) * (1 + FRMC_MAX_GROUP) + sizeof(float) * 3 * k,
                              cudaMemcpyHostToDevice, s->stream));
    prep_proposal_kernel<<<1, FRMC_MAX_GROUP, 0, s->stream>>>(s->d_prop_in, s->d_inv, s->d_atoms, s->d_prop);
    FRMC_LAUNCH_CHECK();
    GridSet gs = make_gridset(s);
    long long want = (s->npad + 255) / 256;
    long long cap = (long long)s->ctx->sm_count * 8;
    int grid = (int)std::max<long long>(1, std::min(want, cap));
#define LAUNCH_DELTA(M) delta_kernel<M><<<grid, 256, 0, s->stream>>>(s->d_atoms, (int)s->npad, s->d_prop, s->L, gs, s->nEl, s->d_overflow)
    cudaEvent_t t0 = timing_begin(s);
    switch (mode) {
        case MODE_IBC: LAUNCH_DELTA(MODE_IBC); break;
        case MODE_ORTHO_FAST: LAUNCH_DELTA(MODE_ORTHO_FAST); break;
        case MODE_TRI_FAST: LAUNCH_DELTA(MODE_TRI_FAST); break;
        case MODE_ORTHO_GEN: LAUNCH_DELTA(MODE_ORTHO_GEN); break;
        default: LAUNCH_DELTA(MODE_TRI_GEN); break;
    }
#undef LAUNCH_DELTA
    FRMC_LAUNCH_CHECK();
    timing_end(s, TIME_DELTA, t0);
    int rc = launch_epilogue(s);
    if (rc) return rc;
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    timing_flush(s);
    for (size_t i = 0; i < s->models.size(); ++i) {
        s->chi2_staged[i] = s->h_chi2[i];
        if (chi2_after) chi2_after[i] = s->h_chi2[i];
    }
    s->state = 1;
    return FRMC_OK;
}

int frmc_accept(frmc_store *s)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->state == 1, FRMC_ESTATE, "no staged proposal to accept");
    FRMC_CUDA(cudaSetDevice(s->dev));
    GridSet gs = make_gridset(s);
    long long cells = 0;
    for (auto &g : s->grids) cells = std::max(cells, 2 * g.dev.cells);
    int grid = (int)std::max<long long>(1, std::min<long long>((cells + 255) / 256, (long long)s->ctx->sm_count * 4));
    cudaEvent_t t0 = timing_begin(s);
    commit_kernel<<<grid, 256, 0, s->stream>>>(gs, s->d_atoms, s->d_prop);
    FRMC_LAUNCH_CHECK();
    timing_end(s, TIME_COMMIT, t0);
    for (auto &m : s->models)
        FRMC_CUDA(cudaMemcpyAsync(m.total_committed, m.dev.total, sizeof(float) * m.dev.n_out, cudaMemcpyDeviceToDevice, s->stream));
    for (int c = 0; c < 3; ++c) { s->lo[c] = std::min(s->lo[c], s->prop_lo[c]); s->hi[c] = std::max(s->hi[c], s->prop_hi[c]); }
    for (size_t i = 0; i < s->models.size(); ++i) s->chi2_committed[i] = s->chi2_staged[i];
    s->state = 0;
    return FRMC_OK;
}

int frmc_reject(frmc_store *s)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->state == 1, FRMC_ESTATE, "no staged proposal to reject");
    FRMC_CUDA(cudaSetDevice(s->dev));
    GridSet gs = make_gridset(s);
    long long cells = 0;
    for (auto &g : s->grids) cells = std::max(cells, 2 * g.dev.cells);
    int grid = (int)std::max<long long>(1, std::min<long long>((cells + 255) / 256, (long long)s->ctx->sm_count * 4));
    cudaEvent_t t0 = timing_begin(s);
    clear_delta_kernel<<<grid, 256, 0, s->stream>>>(gs);
    FRMC_LAUNCH_CHECK();
    timing_end(s, TIME_COMMIT, t0);
    s->state = 0;
    return FRMC_OK;
}

int frmc_export_data(frmc_store *s, int grid, float *hintra, float *hinter)
{
    FRMC_REQUIRE(s && hintra && hinter, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(grid >= 0 && grid < (int)s->grids.size(), FRMC_EINVAL, "unknown grid %d", grid);
    FRMC_CUDA(cudaSetDevice(s->dev));
    GridDev &G = s->grids[grid].dev;
    float *d_out = (float *)ctx_buffer(s->ctx, 5, sizeof(float) * 2 * G.cells);
    if (!d_out) return FRMC_ENOMEM;
    int rc = launch_counts64_to_float(s->stream, G.counts, d_out, 2 * G.cells);
    if (rc) return rc;
    FRMC_CUDA(cudaMemcpyAsync(hintra, d_out, sizeof(float) * G.cells, cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaMemcpyAsync(hinter, d_out + G.cells, sizeof(float) * G.cells, cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    return FRMC_OK;
}

int frmc_export_total(frmc_store *s, int model, int staged, float *out)
{
    FRMC_REQUIRE(s && out, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(model >= 0 && model < (int)s->models.size(), FRMC_EINVAL, "unknown model %d", model);
    FRMC_CUDA(cudaSetDevice(s->dev));
    ModelHost &m = s->models[model];
    FRMC_CUDA(cudaMemcpyAsync(out, staged ? m.dev.total : m.total_committed, sizeof(float) * m.dev.n_out,
                              cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    return FRMC_OK;
}

int frmc_store_set_timing(frmc_store *s, int on)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    s->timing = on != 0;
    if (on) for (int i = 0; i < 4; ++i) { s->kernel_ms[i] = 0; s->kernel_launches[i] = 0; }
    return FRMC_OK;
}

int frmc_store_get_timing(frmc_store *s, int which, double *ms_total, uint64_t *launches)
{
    FRMC_REQUIRE(s && which >= 0 && which < 4, FRMC_EINVAL, "bad timing query");
    FRMC_CUDA(cudaSetDevice(s->dev));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    timing_flush(s);
    if (ms_total) *ms_total = s->kernel_ms[which];
    if (launches) *launches = s->kernel_launches[which];
    return FRMC_OK;
}

uint64_t frmc_store_edge_overflow(frmc_store *s)
{
    if (!s) return 0;
    unsigned long long ov = 0;
    cudaSetDevice(s->dev);
    cudaMemcpyAsync(&ov, s->d_overflow, sizeof(ov), cudaMemcpyDeviceToHost, s->stream);
    cudaStreamSynchronize(s->stream);
    return ov;
}

}  // extern "C"
