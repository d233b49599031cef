// store.cu -- the stateful fast path: device-resident atom store, running integer
// histograms per r-grid, fused per-move delta pass and the constraint-level epilogue
// (G(r) / g(r) / S(Q) / chi^2) on device.
//
// Replaces compute_data / compute_before_move / compute_after_move / accept_move /
// reject_move of PairDistributionConstraint, PairCorrelationConstraint and
// StructureFactorConstraint (Constraints/PairDistributionConstraints.py:1001-1166,
// PairCorrelationConstraints.py:263-392, StructureFactorConstraints.py:933-1096).
//
// Per Metropolis step the device runs TWO kernels (delta pass, fused epilogue) and, after
// the host's decision, one commit-or-clear kernel; the host spins on a pinned completion
// counter instead of synchronising the stream.
#include "common.cuh"
#include "layout.h"
#include "store_view.h"
#include "rng.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <type_traits>
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace frmc {

GridParams make_grid(float rmin, float rmax, float bin, int hs);
int full_hist_launch(cudaStream_t stream, int sm_count, int mode, const float4 *atoms, const uint32_t *orig,
                     int64_t npad, float4 *bbox, const WorkItem *rows, int n_rows, int n_pairs, PairLists &lists,
                     const int32_t *mol_by_orig, uint32_t mol_span,
                     const Lattice &L, const GridParams &g, int nEl, unsigned long long *counts, unsigned long long *stats);
void pack_rows(const std::vector<WorkItem> &rows, std::vector<unsigned char> &blob, int &n_pairs);
int launch_counts64_to_float(cudaStream_t stream, const unsigned long long *counts, float *out, long long cells2);

// ------------------------------------------------------------------ device-side descriptors
// Per grid the store keeps the reference's ORDERED arrays data["intra"], data["inter"]
// (counts, int64, signed) plus the SYMMETRISED totals per unordered element pair
//     tot[sym(a,b)][r] = intra[a,b]+intra[b,a]+inter[a,b]+inter[b,a]   (a != b)
//                      = intra[a,a]+inter[a,a]                          (a == b)
// which is all that reaches G(r) (PairDistributionConstraints.py:867-876).  `stot` is the
// STAGED copy (committed totals + the proposal's signed events): the epilogue reads only
// stot (nEl(nEl+1)/2 cells per bin) instead of 2*nEl^2 ordered cells + deltas.
struct GridDev {
    GridParams g;
    int nsym;                        // nEl*(nEl+1)/2
    int pad;
    long long cells;                 // nEl*nEl*hs
    unsigned long long *counts;      // committed ordered counts [2][cells]  (0 intra, 1 inter)
    int *delta;                      // staged after-minus-before, ordered  [2][cells]
    int *tot;                        // committed symmetrised totals (int32) [nsym][hs]
    int *stot;                       // staged symmetrised totals (tot + proposal, int32) [nsym][hs]
};

struct GridSet {
    int n;
    float t2lo, t2hi;                // union of the grids' d^2 windows (cheap first test)
    int pad;
    GridDev grid[FRMC_MAX_GRIDS];
};

static const int EPI_INLINE_PAIRS = 16;   // pair tables up to this size ride in the kernel parameters

struct ModelDev {
    int kind, grid, n_pairs, n_out, hs, sq_exact;
    float scale;
    const int *psym;                 // [n_pairs] symmetrised pair index of (idi, idj)
    const float *w, *D, *rD, *sv, *pref, *shape, *expv, *wts, *gr2sq;   // rD[p] = RN(1/D[p]) when the 3-op division is proven exact, else NaN
    float *rfun;                     // [hs]    r-space function of the staged state (G(r) or g(r))
    float *total;                    // [n_out] staged model total
    const int *pw_sched;             // [4][pw_leaves] numpy pairwise-sum schedule for n_out terms
    int pw_leaves;
    int nq_pad;                      // row stride of gr2sq on the device (n_out rounded up to 32, zero filled)
    int n_stages;                    // S(Q) ring depth chosen by the host from the shared-memory budget
    int refit;                       // this evaluation refits the scale factor (engine.accepted % frequency == 0; set per launch)
    float sf_min, sf_max;            // clip range of the fitted value (Core/Constraint.py:1392-1393)
    const float *prior;              // [n_out] multiframe prior or NULL: total = prior + mf_weight * total (Core/Constraint.py:1160-1177)
    float mf_weight;
    const float *window;             // [n_window] normalised window function or NULL: total = np.convolve(total, window, "same")
    int n_window;
    // pair table inline (n_pairs <= EPI_INLINE_PAIRS): travels in the kernel parameters, no dependent load
    int i_psym[EPI_INLINE_PAIRS];
    float i_w[EPI_INLINE_PAIRS], i_D[EPI_INLINE_PAIRS], i_rD[EPI_INLINE_PAIRS];
};

struct ModelSet {                    // passed BY VALUE to the epilogue (constant bank: no dependent descriptor load)
    int n;
    int pad;
    ModelDev m[FRMC_MAX_MODELS];
};

__host__ __device__ __forceinline__ int sym_index(int a, int b, int nEl)
{
    if (a > b) { int t = a; a = b; b = t; }
    return a * nEl - (a * (a - 1)) / 2 + (b - a);
}

// ------------------------------------------------------------------ kernels: bookkeeping
// tot, stot <- symmetrised sums of the ordered counts (after compute_data / all-reduce)
// Symmetrised cells are int32: enough for every configuration whose cells stay below 2^31 (the
// fp32 reference is only exact below 2^24 per cell); larger totals raise *too_big.
__global__ void symmetrise_kernel(GridDev G, int nEl, int *__restrict__ too_big)
{
    const int hs = G.g.hs;
    const long long n = (long long)G.nsym * hs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(i / hs), r = (int)(i - (long long)s * hs);
        // invert sym_index: find a with first(a) <= s < first(a+1)
        int a = 0;
        while (a + 1 < nEl && (a + 1) * nEl - ((a + 1) * a) / 2 <= s) ++a;
        const int b = a + (s - (a * nEl - (a * (a - 1)) / 2));
        const long long ab = ((long long)a * nEl + b) * hs + r, ba = ((long long)b * nEl + a) * hs + r;
        long long v = (long long)G.counts[ab] + (long long)G.counts[G.cells + ab];
        if (a != b) v += (long long)G.counts[ba] + (long long)G.counts[G.cells + ba];
        if (v > 0x3FFFFFFFll || v < -0x3FFFFFFFll) *too_big = 1;
        G.tot[i] = (int)v;
        G.stot[i] = (int)v;
    }
}

struct TotalsCopy {
    int n;
    int len[FRMC_MAX_MODELS];
    const float *src[FRMC_MAX_MODELS];
    float *dst[FRMC_MAX_MODELS];
};

// accept: fold the staged deltas into the committed state, clear them, move the atoms and
// promote the staged model totals to committed
__global__ void commit_kernel(GridSet gs, float4 *__restrict__ atoms, const Proposal *__restrict__ prop, TotalsCopy tc)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &G = gs.grid[gi];
        for (long long c = tid; c < 2 * G.cells; c += stride) {
            const int d = G.delta[c];
            if (d) { G.counts[c] = (unsigned long long)((long long)G.counts[c] + d); G.delta[c] = 0; }
        }
        const long long ns = (long long)G.nsym * G.g.hs;
        for (long long c = tid; c < ns; c += stride) G.tot[c] = G.stot[c];
    }
    if (tid < prop->k) atoms[prop->pos[tid]] = prop->newc[tid];
    for (int m = 0; m < tc.n; ++m)
        for (long long i = tid; i < tc.len[m]; i += stride) tc.dst[m][i] = tc.src[m][i];
}

// reject: clear the staged deltas
__global__ void clear_delta_kernel(GridSet gs)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &G = gs.grid[gi];
        for (long long c = tid; c < 2 * G.cells; c += stride)
            if (G.delta[c]) G.delta[c] = 0;
        const long long ns = (long long)G.nsym * G.g.hs;
        for (long long c = tid; c < ns; c += stride) G.stot[c] = G.tot[c];
    }
}

// ------------------------------------------------------------------ kernels: delta pass
__device__ __forceinline__ void delta_hit(float d2, int sign, int same, int slab, int sym, const GridSet &gs, int nEl,
                                          unsigned long long &ov)
{
#pragma unroll 1
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &G = gs.grid[gi];
        if (in_range(d2, G.g)) {
            int b = bin_index(d2, G.g);
            if (b < G.g.hs) {
                atomicAdd(&G.delta[(same ? 0 : G.cells) + (long long)slab * G.g.hs + b], sign);
                atomicAdd(&G.stot[(long long)sym * G.g.hs + b], sign);
            } else {
                ++ov;
                // reference-compatible spill of the unchecked write: flat index runs into the next slab
                const long long flat = (long long)slab * G.g.hs + b;
                if (G.g.spill && flat < G.cells) {
                    const int s2 = (int)(flat / G.g.hs), b2 = (int)(flat - (long long)s2 * G.g.hs);
                    atomicAdd(&G.delta[(same ? 0 : G.cells) + flat], sign);
                    atomicAdd(&G.stot[(long long)sym_index(s2 / nEl, s2 % nEl, nEl) * G.g.hs + b2], sign);
                }
            }
        }
    }
}

// One streaming pass over the whole store: every atom record is read ONCE (coalesced
// 16-byte loads, DELTA_UNROLL independent loads in flight per thread) and tested against the
// old and the new position of every group atom; the signed events (-1 old, +1 new) land in
// the int32 delta histograms of every grid.  Pairs inside the group are handled by block 0
// with the reference's M-F convention (PairDistributionConstraints.py:1053-1078: the pair
// (t,u), u earlier in the index list, survives once in slab [el_t, el_u]).
// compute_before_move + compute_after_move = this one launch.  Algorithmic traffic: 16 B/atom.
static const int DELTA_UNROLL = 4;

struct DeltaShared {
    float4 sOld[FRMC_MAX_GROUP];
    float4 sNew[FRMC_MAX_GROUP];
    int sPos[FRMC_MAX_GROUP];
};

// body of the delta pass, shared by the stand-alone kernel and the fused propose kernel
template <int MODE>
__device__ __forceinline__ void delta_body(DeltaShared &sh, const float4 *__restrict__ atoms, int npad, const ProposalIn &in,
                                           Proposal *__restrict__ prop, const Lattice &L, const GridSet &gs, int nEl,
                                           unsigned long long *__restrict__ overflow)
{
    float4 *sOld = sh.sOld, *sNew = sh.sNew;
    int *sPos = sh.sPos;
    const int k = in.k;
    unsigned long long ov = 0;
    const int T = gridDim.x * blockDim.x;
    // double-buffered: the DELTA_UNROLL loads of the next batch are in flight while this batch is processed.  The first
    // batch is requested BEFORE the moved atoms are staged: the two L2 round trips overlap.
    const float4 padrec = make_float4(0.f, 0.f, 0.f, __uint_as_float(PAD_META));
    float4 nxt[DELTA_UNROLL];
    {
        const int p0 = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
        for (int u = 0; u < DELTA_UNROLL; ++u) { const int p = p0 + u * T; nxt[u] = (p < npad) ? __ldcg(atoms + p) : padrec; }
    }
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const int p = in.pos[t];
        const float4 o = __ldcg(atoms + p);
        const float4 nw = make_float4(in.moved[3 * t], in.moved[3 * t + 1], in.moved[3 * t + 2], o.w);
        sOld[t] = o; sNew[t] = nw; sPos[t] = p;
        if (blockIdx.x == 0) { prop->pos[t] = p; prop->newc[t] = nw; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) prop->k = k;
    __syncthreads();
    for (int p0 = blockIdx.x * blockDim.x + threadIdx.x; p0 < npad; p0 += DELTA_UNROLL * T) {
        float4 a[DELTA_UNROLL];
#pragma unroll
        for (int u = 0; u < DELTA_UNROLL; ++u) a[u] = nxt[u];
        {
            const int q0 = p0 + DELTA_UNROLL * T;
#pragma unroll
            for (int u = 0; u < DELTA_UNROLL; ++u) { const int p = q0 + u * T; nxt[u] = (p < npad) ? __ldcg(atoms + p) : padrec; }
        }
#pragma unroll
        for (int u = 0; u < DELTA_UNROLL; ++u) {
            const int p = p0 + u * T;
            const uint32_t mj = __float_as_uint(a[u].w);
            if (mj == PAD_META) continue;
            bool ingroup = false;
            for (int t = 0; t < k; ++t) ingroup |= (sPos[t] == p);
            if (ingroup) continue;
            const int ej = mj & 0xFF;
            for (int t = 0; t < k; ++t) {
                const float4 o = sOld[t], nw = sNew[t];
                const uint32_t mt = __float_as_uint(o.w);
                const float d2o = dist2<MODE>(o.x, o.y, o.z, a[u].x, a[u].y, a[u].z, L);
                const float d2n = dist2<MODE>(nw.x, nw.y, nw.z, a[u].x, a[u].y, a[u].z, L);
                const bool ho = (d2o >= gs.t2lo) && (d2o < gs.t2hi);
                const bool hn = (d2n >= gs.t2lo) && (d2n < gs.t2hi);
                if (ho || hn) {
                    const int same = (mt >> 8) == (mj >> 8);
                    const int et = (int)(mt & 0xFF);
                    const int slab = et * nEl + ej;
                    const int sym = sym_index(et, ej, nEl);
                    if (ho) delta_hit(d2o, -1, same, slab, sym, gs, nEl, ov);
                    if (hn) delta_hit(d2n, +1, same, slab, sym, gs, nEl, ov);
                }
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
            const int et = (int)(mt & 0xFF), eu = (int)(mu & 0xFF);
            const int slab = et * nEl + eu;
            const int sym = sym_index(et, eu, nEl);
            const float d2o = dist2<MODE>(ot.x, ot.y, ot.z, ou.x, ou.y, ou.z, L);
            const float d2n = dist2<MODE>(nt.x, nt.y, nt.z, nu.x, nu.y, nu.z, L);
            if ((d2o >= gs.t2lo) && (d2o < gs.t2hi)) delta_hit(d2o, -1, same, slab, sym, gs, nEl, ov);
            if ((d2n >= gs.t2lo) && (d2n < gs.t2hi)) delta_hit(d2n, +1, same, slab, sym, gs, nEl, ov);
        }
    }
    if (ov) atomicAdd(overflow, ov);
}

template <int MODE>
__global__ void __launch_bounds__(256)
delta_kernel(const float4 *__restrict__ atoms, int npad, const ProposalIn in, Proposal *__restrict__ prop,
             Lattice L, GridSet gs, int nEl, unsigned long long *__restrict__ overflow, long long *__restrict__ stamps)
{
    __shared__ DeltaShared sh;
    if (stamps && threadIdx.x == 0) {       // debug timeline (globaltimer ns): first start / last end over all CTAs
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMin(reinterpret_cast<unsigned long long *>(stamps + 120), gt);
    }
    delta_body<MODE>(sh, atoms, npad, in, prop, L, gs, nEl, overflow);
    if (stamps && threadIdx.x == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long *>(stamps + 121), gt);
    }
}

// ------------------------------------------------------------------ kernels: fused epilogue
// numpy's pairwise float32 summation (numpy/_core/src/umath/loops_utils.h.src,
// FLOAT_pairwise_sum): blocks of <=128 summed with 8 interleaved accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), remainder added sequentially, larger arrays split
// at n/2 rounded down to a multiple of 8.  Reproduced exactly so chi^2 equals
// np.add.reduce(w*(exp-model)**2) bit for bit (PairDistributionConstraints.py:833-838).
static const int PW_MAX_LEAVES = 512;
static const int EPI_THREADS = 512;
static const int EPI_MAX_PAIRS = FRMC_MAX_ELEMENTS * (FRMC_MAX_ELEMENTS + 1) / 2;

// The recursion tree depends only on n, so the host precomputes it once per model
// (pairwise_schedule): the leaves (offset, length) left to right, and the post-order list of
// combine steps val[dst] += val[src] over the leaf sums (n_leaves - 1 steps, result in val[0]).
struct PairwiseScratch {
    int leaf_off[PW_MAX_LEAVES];
    int leaf_len[PW_MAX_LEAVES];
    int op_dst[PW_MAX_LEAVES];
    int op_src[PW_MAX_LEAVES];
    float leaf_sum[PW_MAX_LEAVES];
};

// whole CTA (EPI_THREADS threads, all must call); v[0..n) in shared memory; the schedule
// (leaves + combine steps) is already staged in ps; result valid in thread 0
__device__ float block_pairwise_sum(const float *v, int nl, PairwiseScratch &ps)
{
    const int group = threadIdx.x >> 3, lane8 = threadIdx.x & 7;
    for (int base = 0; base < nl; base += EPI_THREADS / 8) {
        const int l = base + group;
        const bool live = l < nl;
        const int off = live ? ps.leaf_off[l] : 0, len = live ? ps.leaf_len[l] : 0;
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
            ps.leaf_sum[l] = res;
        }
    }
    __syncthreads();
    float out = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nl - 1; ++i) ps.leaf_sum[ps.op_dst[i]] = __fadd_rn(ps.leaf_sum[ps.op_dst[i]], ps.leaf_sum[ps.op_src[i]]);
        out = ps.leaf_sum[0];
    }
    return out;
}

static const int SQ_ROWS = 64;       // matrix rows per cp.async stage (8 KB per 32-column slab)
static const int SQ_MAX_STAGES = 26;  // ring stages (8 KB each); a whole [hs x 32] slab stays resident when hs <= 64*stages

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > 50000000u) __trap();     // never hang the GPU on a lost transaction
}

// (w*n)/D with IEEE rounding.  D is a per-pair constant, so the quotient is formed from the
// correctly rounded reciprocal with one residual correction (Markstein): q0 = x*rD,
// r = fma(-q0, D, x) (exact), q = fma(r, rD, q0).  For every pair this shortcut is CHECKED
// against __fdiv_rn over the whole count range at model registration (validate_fastdiv_kernel);
// a pair that fails any count keeps the full division (rD = NaN).
__device__ __forceinline__ float div_by_const(float x, float D, float rD)
{
    const float q0 = __fmul_rn(x, rD);
    const float r = __fmaf_rn(-q0, D, x);
    return __fmaf_rn(r, rD, q0);
}

// ok[p] stays 1 iff div_by_const(w*n, D, rD) == __fdiv_rn(w*n, D) for every integer count n in
// [0, 2^22] and for 2^22 further counts spread up to 2^30 (the int32 running totals' range).
__global__ void validate_fastdiv_kernel(const float *__restrict__ w, const float *__restrict__ D, int n_pairs, int *__restrict__ ok,
                                        float *__restrict__ rD_out)
{
    const int p = blockIdx.y;
    if (p >= n_pairs) return;
    const float wp = w[p], Dp = D[p], rD = __frcp_rn(Dp);
    if (blockIdx.x == 0 && threadIdx.x == 0) rD_out[p] = rD;
    bool good = true;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ll << 23); i += (long long)gridDim.x * blockDim.x) {
        const long long n = (i < (1ll << 22)) ? i : ((i - (1ll << 22)) * 255 + (1ll << 22));
        const float x = __fmul_rn(wp, (float)(int)n);
        if (__float_as_uint(div_by_const(x, Dp, rD)) != __float_as_uint(__fdiv_rn(x, Dp))) good = false;
    }
    if (!good) ok[p] = 0;
}

// ONE launch for every model: grid = (max Q slices, n_models), EPI_THREADS threads; the model
// descriptors travel by value in the kernel parameters.
//
//  1. r-space function of the staged state for ALL bins, redundantly per CTA (it only needs the
//     nsym x hs staged symmetrised totals, L2 resident), mirroring the numpy expressions of
//     __get_total_Gr / __get_total_gr / __get_total_Sq operation by operation in fp32
//     (numpy >= 2 scalar promotion: every scalar is fp32):
//        for pair in sorted pairs:  Gr += (wij*nij)/Dij          (PairDistributionConstraints.py:855-876)
//        Gr /= shellVolumes                                      (:878)
//        Gr  = prefactor*(Gr-1)                                  (:881)   [PDF, SQ, RSQ]
//        Gr -= shape ; Gr *= scale (when != 1)                   (:883-888) [PDF]
//        PCF: gr -= shape; if scale != 1: gr = 1 + (prefactor*(gr-1)*scale)/prefactor  (PairCorrelationConstraints.py:153-163)
//     All loads of a thread's bins are issued before any arithmetic (one L2 round trip).
//  2. PDF/PCF (one CTA): chi^2 with numpy's pairwise order, publish.
//  3. SQ/RSQ (CTA x owns Q columns [32x, 32x+32)): S(Q_m) = sum_r G(r)*M[r,m] (+1), r in index
//     order, fp32 multiply then fp32 add, no FMA: bit-identical to
//     np.sum(Gr.reshape((-1,1))*Gr2SqMatrix, axis=0) (StructureFactorConstraints.py:772-773),
//     which accumulates rows sequentially.  The chain of hs dependent FADDs (4 cycles each) is
//     the floor, ~2 us at hs=1000.  The matrix is stored pre-tiled ([Q slab][r/4][lane][r%4],
//     zero padded to 64 rows), so a 64-row chunk of a slab is 8 KB contiguous: lane 0 fetches
//     the whole slab with TMA bulk copies (cp.async.bulk + one mbarrier per chunk) at kernel
//     start, up to 26 chunks = 1664 rows resident in shared memory (longer slabs reuse the
//     stages as a ring); warp 0 reads four rows per LDS.128 and forms products 16 rows ahead.
//     The last CTA to finish (ticket) computes chi^2.
//  4. publish: chi2 to pinned host memory, system fence, then the per-model launch counter the
//     host spins on.
static const unsigned SQ_CHUNK_BYTES = SQ_ROWS * 32 * sizeof(float);

// 16-byte shared load the compiler may not sink towards its use (software-pipelined S(Q) loop)
__device__ __forceinline__ float4 lds128(const float4 *p)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

// lane 0 of the consumer warp only
__device__ __forceinline__ void sq_issue(const ModelDev &M, float *ring, unsigned long long *mbar, int slab, int chunk, int n_chunks)
{
    if (chunk < n_chunks) {
        const int stage = chunk % M.n_stages;
        const float *src = M.gr2sq + ((size_t)slab * n_chunks + chunk) * (SQ_ROWS * 32);
        mbar_expect_tx(&mbar[stage], SQ_CHUNK_BYTES);
        bulk_g2s(ring + stage * (SQ_ROWS * 32), src, SQ_CHUNK_BYTES, &mbar[stage]);
    }
}

struct EpiShared {
    int s_psym[EPI_MAX_PAIRS + 16];
    float s_w[EPI_MAX_PAIRS + 16], s_D[EPI_MAX_PAIRS + 16], s_rD[EPI_MAX_PAIRS + 16];
    PairwiseScratch ps;
    int s_last;
    float s_sf;                      // scale factor of this evaluation (fitted or the committed one)
    unsigned long long mbar[SQ_MAX_STAGES];
};

// element i of np.convolve(a, v, "same") for len(a) = n >= len(v) = nw (sequential fp32 accumulation; numpy's
// correlate uses its dot kernel, whose summation order is not specified: this branch is within 1e-6, not bit-exact)
template <typename FA>
__device__ __forceinline__ float convolve_same(FA a, int n, const float *__restrict__ v, int nw, int i)
{
    const int nf = i + (nw - 1) / 2;
    const int m0 = max(0, nf - (nw - 1)), m1 = min(n - 1, nf);
    float acc = 0.0f;
    for (int m = m0; m <= m1; ++m) acc = __fadd_rn(acc, __fmul_rn(a(m), v[nf - m]));
    return acc;
}

// ExperimentalConstraint.fit_scale_factor (Core/Constraint.py:1363-1395): SF = sum(w*M*E) / sum(M**2) in
// numpy's fp32 pairwise order, clipped.  mv/ev = model / experimental value of element i as the calling
// constraint hands them over (G(r); 4*pi*r*rho0*(g-1); S(Q)-1).  Whole CTA; result in es.s_sf.
template <typename FM, typename FE>
__device__ __forceinline__ void fit_scale_factor(EpiShared &es, float *sT, const ModelDev &M, int n, FM mv, FE ev)
{
    for (int i = threadIdx.x; i < n; i += EPI_THREADS) {
        const float m = mv(i);
        sT[i] = M.wts ? __fmul_rn(__fmul_rn(M.wts[i], m), ev(i)) : __fmul_rn(m, ev(i));
    }
    __syncthreads();
    const float s1 = block_pairwise_sum(sT, M.pw_leaves, es.ps);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += EPI_THREADS) { const float m = mv(i); sT[i] = __fmul_rn(m, m); }
    __syncthreads();
    const float s2 = block_pairwise_sum(sT, M.pw_leaves, es.ps);
    if (threadIdx.x == 0) {
        float sf = __fdiv_rn(s1, s2);
        if (M.sf_min > sf) sf = M.sf_min;          // max(SF, minimum): Python keeps SF unless minimum > SF
        if (M.sf_max < sf) sf = M.sf_max;          // min(SF, maximum)
        es.s_sf = sf;
    }
    __syncthreads();
}

#define EPI_STAMP(i) do { if (stamps && threadIdx.x == 0 && slab == 0) { stamps[m * 8 + (i)] = clock64(); \
        unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); stamps[64 + m * 8 + (i)] = (long long)gt_; } } while (0)

// part 1 of an epilogue CTA (model M, Q slab `slab`): start the TMA stream of the matrix slab.
// Called first thing in the kernel so the copies overlap everything that precedes the S(Q) loop.
__device__ __forceinline__ void epilogue_prefetch(EpiShared &es, float *epi_smem, const ModelDev &M, int slab)
{
    const bool is_sq = (M.kind == FRMC_KIND_SQ || M.kind == FRMC_KIND_RSQ);
    const int hs = M.hs, nq = M.n_out;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int hs_pad = (hs + SQ_ROWS - 1) / SQ_ROWS * SQ_ROWS;
    const int nst = M.n_stages;
    const int q0 = slab * 32;
    float *ring = epi_smem;                              // first: keeps the TMA destinations 128-byte aligned
    float *sG = epi_smem + (is_sq ? nst * SQ_ROWS * 32 : 0);
    float *sT = sG + hs_pad;
    const int n_chunks = (hs + SQ_ROWS - 1) / SQ_ROWS;
    const bool resident = n_chunks <= nst;
    unsigned long long *mbar = es.mbar;
    (void)lane; (void)wrp; (void)q0; (void)ring; (void)sT; (void)nq; (void)resident; (void)mbar;
    // start streaming the matrix slab before anything else: it overlaps the G(r) phase.
    // Resident case (whole slab fits the ring): ONE mbarrier, a few large bulk copies.
    if (is_sq && wrp == 0) {
        if (lane == 0) {
            for (int st = 0; st < (resident ? 1 : nst); ++st) mbar_init(&mbar[st], 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            if (resident) {
                const unsigned total = (unsigned)n_chunks * SQ_CHUNK_BYTES;
                const float *src = M.gr2sq + (size_t)slab * n_chunks * (SQ_ROWS * 32);
                mbar_expect_tx(&mbar[0], total);
                for (unsigned off = 0; off < total; off += 4 * SQ_CHUNK_BYTES) {
                    const unsigned bytes = min(4 * SQ_CHUNK_BYTES, total - off);
                    bulk_g2s(reinterpret_cast<char *>(ring) + off, reinterpret_cast<const char *>(src) + off, bytes, &mbar[0]);
                }
            } else {
                for (int c = 0; c < nst; ++c) sq_issue(M, ring, mbar, slab, c, n_chunks);
            }
        }
        __syncwarp();
    }

}

// Where a BATCH evaluation (batch_kernel: several proposals in flight) reads its counts and leaves its results:
// the counts are the COMMITTED symmetrised totals plus the proposal's own symmetrised delta, the model total and
// chi^2 go to per-proposal buffers, and nothing is published to the host (the kernel decides itself).
static const int EPI_MAX_ADDS = 12;
struct EpiOut {
    const int *base;         // [nsym][hs] symmetrised totals the evaluation starts from (a fully visible copy)
    const int *adds[EPI_MAX_ADDS];   // [nsym][hs] symmetrised deltas on this model's grid: the proposal's own, then those
    unsigned int amask[EPI_MAX_ADDS];// of the proposals assumed accepted and of the acceptances not yet folded into base;
    int n_adds;              // amask bit s: row s can be non-zero (rows of untouched element pairs are not read)
    float *total;            // [n_out]    model total of this evaluation
    float *res;              // [2*FRMC_MAX_MODELS] chi2 per model, then the scale factor each evaluation used
    float *terms;            // S(Q) models: [n_out] weighted squared residuals; when set the slab CTAs stop there and
                             // every CTA forms chi2 itself after the grid barrier (no ticket, no last-CTA stage)
    int warm;                // this CTA has run this (model, slab) before in this launch: tables and schedule are staged
    const unsigned int *ev;  // on-the-fly pair corrections of this node on this model's grid: (row << 16 | bin), bit 31 = minus one
    const unsigned int *ev_bits;   // 1024-bit filter on (bin & 1023): a thread scans the list only when one of its bins is marked
    int n_ev;
    int refit;               // this evaluation refits the scale factor (the engine's accepted count AT THIS NODE % frequency == 0)
    float scale;             // the model's committed scale factor at this node (acceptances inside the launch may have changed it)
};

// part 2: r-space function, chi^2 / S(Q) slice, ticket, publish (steps 1-4 above)
template <bool BATCH>
__device__ __forceinline__ void epilogue_run(EpiShared &es, float *epi_smem, const ModelSet &ms, const GridSet &gs, int m, int slab,
                                             float *__restrict__ chi2_out, unsigned int *__restrict__ dev_seq,
                                             volatile unsigned int *__restrict__ host_seq, unsigned int *__restrict__ tickets,
                                             long long *__restrict__ stamps, const EpiOut &eo)
{
    const ModelDev &M = ms.m[m];
    float *const total = BATCH ? eo.total : M.total;
    const int mrefit = BATCH ? eo.refit : M.refit;           // per launch on the one-proposal paths, per node in a batch
    const float mscale = BATCH ? eo.scale : M.scale;
    const bool is_sq = (M.kind == FRMC_KIND_SQ || M.kind == FRMC_KIND_RSQ);
    const int hs = M.hs, nq = M.n_out;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int hs_pad = (hs + SQ_ROWS - 1) / SQ_ROWS * SQ_ROWS;
    const int nst = M.n_stages;
    const int q0 = slab * 32;
    float *ring = epi_smem;                              // first: keeps the TMA destinations 128-byte aligned
    float *sG = epi_smem + (is_sq ? nst * SQ_ROWS * 32 : 0);
    float *sT = sG + hs_pad;
    const int n_chunks = (hs + SQ_ROWS - 1) / SQ_ROWS;
    const bool resident = n_chunks <= nst;
    unsigned long long *mbar = es.mbar;
    (void)lane; (void)wrp; (void)q0; (void)ring; (void)sT; (void)nq; (void)resident; (void)mbar;
    const int nblk = is_sq ? (nq + 31) / 32 : 1;
    int *s_psym = es.s_psym;
    float *s_w = es.s_w, *s_D = es.s_D, *s_rD = es.s_rD;
    PairwiseScratch &ps = es.ps;
    int &s_last = es.s_last;
    // pair table, padded to a multiple of 16 with weight-0 entries: (0*n)/1 = 0 and acc+0 == acc
    const int np = M.n_pairs;
    const int np_pad = (np + 15) / 16 * 16;
    const bool inline_pairs = np <= EPI_INLINE_PAIRS;
    const bool cold = !BATCH || !eo.warm;
    if (cold)
    for (int p = tid; p < np_pad; p += EPI_THREADS) {
        const bool real = p < np;
        if (inline_pairs) {
            s_psym[p] = real ? M.i_psym[p] : 0; s_w[p] = real ? M.i_w[p] : 0.0f;
            s_D[p] = real ? M.i_D[p] : 1.0f; s_rD[p] = real ? M.i_rD[p] : 1.0f;
        } else {
            s_psym[p] = real ? M.psym[p] : 0; s_w[p] = real ? M.w[p] : 0.0f;
            s_D[p] = real ? M.D[p] : 1.0f; s_rD[p] = real ? M.rD[p] : 1.0f;
        }
    }
    if (cold) {
        for (int r = hs + tid; r < hs_pad; r += EPI_THREADS) sG[r] = 0.0f;
        __syncthreads();
    }
    EPI_STAMP(1);

    // ---- 1. r-space function: EPI_BINS bins per thread per round, every load issued before any arithmetic
    {
        const int *__restrict__ stot = BATCH ? eo.base : gs.grid[M.grid].stot;
        constexpr int EPI_BINS = 2;
        const bool defer = !is_sq && (mrefit || M.prior || M.window);   // scale, prior, window applied in stage 1b
        // bin r from the sum over the pair terms (acc): /shellVolumes, prefactor, shape, scale
        auto finish_bin = [&](int r, float accv, float svr, float prf, float shp) {
            float a = __fdiv_rn(accv, svr);
            float out;
            if (M.kind == FRMC_KIND_PCF) {
                out = a;
                if (M.shape) out = __fsub_rn(out, shp);
                if (!defer && mscale != 1.0f) {
                    float Gr = __fmul_rn(prf, __fsub_rn(out, 1.0f));
                    Gr = __fmul_rn(Gr, mscale);
                    out = __fadd_rn(1.0f, __fdiv_rn(Gr, prf));
                }
            } else {
                out = __fmul_rn(prf, __fsub_rn(a, 1.0f));
                if (M.kind == FRMC_KIND_PDF) {
                    if (M.shape) out = __fsub_rn(out, shp);
                    if (!defer && mscale != 1.0f) out = __fmul_rn(out, mscale);
                }
            }
            sG[r] = out;
            if (slab == 0 && !defer) {
                if (!BATCH) M.rfun[r] = out;
                if (!is_sq) total[r] = out;
            }
        };
        // the pair loop is unrolled PB pairs at a time; models with few element pairs (two elements: three pairs) take
        // the narrow variant instead of loading twelve padding rows per bin
        auto bins = [&](auto pb_tag) {
        constexpr int PB = decltype(pb_tag)::value;
        const int np_padx = (np + PB - 1) / PB * PB;
        if (BATCH && (hs & 1) == 0) {
            // batch, even histogram size: two consecutive bins per thread through 8-byte loads (half the load and
            // address instructions of the scalar loop; rows start on 8-byte boundaries because hs is even)
            for (int r0 = 2 * tid; r0 < hs; r0 += 2 * EPI_THREADS) {
                float acc0 = 0.0f, acc1 = 0.0f;
                const float2 svr = *reinterpret_cast<const float2 *>(M.sv + r0), prf = *reinterpret_cast<const float2 *>(M.pref + r0);
                const float2 shp = M.shape ? *reinterpret_cast<const float2 *>(M.shape + r0) : make_float2(0.f, 0.f);
                for (int p0 = 0; p0 < ((M.sq_exact & 4) ? 0 : np_padx); p0 += PB) {
                    int2 c[PB];
#pragma unroll
                    for (int u = 0; u < PB; ++u) c[u] = __ldcg(reinterpret_cast<const int2 *>(stot + s_psym[p0 + u] * hs + r0));
                    for (int x = 0; x < eo.n_adds; ++x) {       // block-uniform
                        const int *ap = eo.adds[x];
                        const unsigned int mk = eo.amask[x];
                        int2 e[PB];
#pragma unroll
                        for (int u = 0; u < PB; ++u)
                            e[u] = ((mk >> (s_psym[p0 + u] & 31)) & 1u) ? __ldcg(reinterpret_cast<const int2 *>(ap + s_psym[p0 + u] * hs + r0))
                                                                        : make_int2(0, 0);
#pragma unroll
                        for (int u = 0; u < PB; ++u) { c[u].x += e[u].x; c[u].y += e[u].y; }
                    }
                    // on-the-fly pair corrections (CorrEv): the few events of this node, filtered by bin
                    if (eo.n_ev && (((eo.ev_bits[(r0 & 1023) >> 5] >> (r0 & 31)) | (eo.ev_bits[((r0 + 1) & 1023) >> 5] >> ((r0 + 1) & 31))) & 1u)) {
                        for (int x = 0; x < eo.n_ev; ++x) {
                            const unsigned int evx = eo.ev[x];
                            const int eb = (int)(evx & 0xFFFFu), erow = (int)((evx >> 16) & 0x7FFFu), es_ = (evx >> 31) ? -1 : 1;
                            if ((eb & ~1) != r0) continue;
#pragma unroll
                            for (int u = 0; u < PB; ++u)
                                if (s_psym[p0 + u] == erow && s_w[p0 + u] != 0.0f) { if (eb & 1) c[u].y += es_; else c[u].x += es_; }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PB; ++u) {
                        const float w = s_w[p0 + u], D = s_D[p0 + u], rD = s_rD[p0 + u];
                        if (rD == rD) {                       // block-uniform: proven 3-op exact division
                            acc0 = __fadd_rn(acc0, div_by_const(__fmul_rn(w, (float)c[u].x), D, rD));
                            acc1 = __fadd_rn(acc1, div_by_const(__fmul_rn(w, (float)c[u].y), D, rD));
                        } else {
                            acc0 = __fadd_rn(acc0, __fdiv_rn(__fmul_rn(w, (float)c[u].x), D));
                            acc1 = __fadd_rn(acc1, __fdiv_rn(__fmul_rn(w, (float)c[u].y), D));
                        }
                    }
                }
                finish_bin(r0, acc0, svr.x, prf.x, shp.x);
                finish_bin(r0 + 1, acc1, svr.y, prf.y, shp.y);
            }
        } else
        for (int rb = tid; rb < hs; rb += EPI_BINS * EPI_THREADS) {
            float acc[EPI_BINS], svr[EPI_BINS], prf[EPI_BINS], shp[EPI_BINS];
#pragma unroll
            for (int j = 0; j < EPI_BINS; ++j) {
                const int r = min(rb + j * EPI_THREADS, hs - 1);
                acc[j] = 0.0f;
                svr[j] = M.sv[r]; prf[j] = M.pref[r]; shp[j] = M.shape ? M.shape[r] : 0.0f;
            }
            for (int p0 = 0; p0 < ((M.sq_exact & 4) ? 0 : np_padx); p0 += PB) {
                int c[EPI_BINS][PB];
#pragma unroll
                for (int j = 0; j < EPI_BINS; ++j) {
                    const int r = min(rb + j * EPI_THREADS, hs - 1);
#pragma unroll
                    for (int u = 0; u < PB; ++u) c[j][u] = __ldcg(stot + (long long)s_psym[p0 + u] * hs + r);
                    if (BATCH) {
                        for (int x = 0; x < eo.n_adds; ++x) {       // block-uniform
                            const int *ap = eo.adds[x];
                            const unsigned int mk = eo.amask[x];
                            int e[PB];
#pragma unroll
                            for (int u = 0; u < PB; ++u)
                                e[u] = ((mk >> (s_psym[p0 + u] & 31)) & 1u) ? __ldcg(ap + (long long)s_psym[p0 + u] * hs + r) : 0;
#pragma unroll
                            for (int u = 0; u < PB; ++u) c[j][u] += e[u];
                        }
                        if (eo.n_ev && ((eo.ev_bits[(r & 1023) >> 5] >> (r & 31)) & 1u)) {      // on-the-fly pair corrections
                            for (int x = 0; x < eo.n_ev; ++x) {
                                const unsigned int evx = eo.ev[x];
                                const int eb = (int)(evx & 0xFFFFu), erow = (int)((evx >> 16) & 0x7FFFu), es_ = (evx >> 31) ? -1 : 1;
                                if (eb != r) continue;
#pragma unroll
                                for (int u = 0; u < PB; ++u)
                                    if (s_psym[p0 + u] == erow && s_w[p0 + u] != 0.0f) c[j][u] += es_;
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const float w = s_w[p0 + u], D = s_D[p0 + u], rD = s_rD[p0 + u];
                    if (rD == rD) {                       // block-uniform: proven 3-op exact division
#pragma unroll
                        for (int j = 0; j < EPI_BINS; ++j)
                            acc[j] = __fadd_rn(acc[j], div_by_const(__fmul_rn(w, (float)c[j][u]), D, rD));
                    } else {
#pragma unroll
                        for (int j = 0; j < EPI_BINS; ++j)
                            acc[j] = __fadd_rn(acc[j], __fdiv_rn(__fmul_rn(w, (float)c[j][u]), D));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < EPI_BINS; ++j) {
                const int r = rb + j * EPI_THREADS;
                if (r >= hs) continue;
                finish_bin(r, acc[j], svr[j], prf[j], shp[j]);
            }
        }
        };
        if (np <= 4) bins(std::integral_constant<int, 4>{});
        else bins(std::integral_constant<int, 16>{});
    }
    {   // chi^2 summation schedule (needed at the very end): fetched here, its latency hides behind the barrier
        const int nl = M.pw_leaves;
        if (cold)
        for (int i = tid; i < nl; i += EPI_THREADS) {
            ps.leaf_off[i] = M.pw_sched[i]; ps.leaf_len[i] = M.pw_sched[nl + i];
            ps.op_dst[i] = M.pw_sched[2 * nl + i]; ps.op_src[i] = M.pw_sched[3 * nl + i];
        }
    }
    __syncthreads();
    EPI_STAMP(2);

    float chi2 = 0.f;
    if (tid == 0) es.s_sf = mscale;
    if (!is_sq) {
        if (mrefit || M.prior || M.window) {
            // ---- 1b. scale-factor refit on the unscaled function (PairDistributionConstraints.py:886-888;
            //          PairCorrelationConstraints.py:171-184 fits on G(r) = 4 pi r rho0 (g - 1)), then scale,
            //          multiframe prior (:890) and window convolution (:892-893)
            float sf = mscale;
            if (mrefit) {
                if (M.kind == FRMC_KIND_PCF)
                    fit_scale_factor(es, sT, M, hs, [&](int i) { return __fmul_rn(M.pref[i], __fsub_rn(sG[i], 1.0f)); },
                                     [&](int i) { return __fmul_rn(M.pref[i], __fsub_rn(M.expv[i], 1.0f)); });
                else
                    fit_scale_factor(es, sT, M, hs, [&](int i) { return sG[i]; }, [&](int i) { return M.expv[i]; });
                sf = es.s_sf;
            }
            for (int i = tid; i < hs; i += EPI_THREADS) {
                float out = sG[i];
                if (sf != 1.0f) {
                    if (M.kind == FRMC_KIND_PCF) {
                        float Gr = __fmul_rn(M.pref[i], __fsub_rn(out, 1.0f));
                        Gr = __fmul_rn(Gr, sf);
                        out = __fadd_rn(1.0f, __fdiv_rn(Gr, M.pref[i]));
                    } else {
                        out = __fmul_rn(out, sf);
                    }
                }
                if (M.prior) out = __fadd_rn(M.prior[i], __fmul_rn(M.mf_weight, out));
                sG[i] = out;
            }
            __syncthreads();
            if (M.window) {
                for (int i = tid; i < hs; i += EPI_THREADS) sT[i] = convolve_same([&](int m) { return sG[m]; }, hs, M.window, M.n_window, i);
                __syncthreads();
                for (int i = tid; i < hs; i += EPI_THREADS) sG[i] = sT[i];
                __syncthreads();
            }
            for (int i = tid; i < hs; i += EPI_THREADS) { if (!BATCH) M.rfun[i] = sG[i]; total[i] = sG[i]; }
            __syncthreads();
        }
        // ---- 2. chi^2 of an r-space model
        for (int i = tid; i < hs; i += EPI_THREADS) {
            float d = __fsub_rn(M.expv[i], sG[i]);
            float t = __fmul_rn(d, d);
            if (M.wts) t = __fmul_rn(M.wts[i], t);
            sT[i] = t;
        }
        __syncthreads();
        chi2 = block_pairwise_sum(sT, M.pw_leaves, ps);
    } else {
        // ---- 3. S(Q) slice: warp 0 consumes the slab chunk by chunk (refilling the ring only when the slab does not fit)
        if (wrp == 0) {
            float acc = 0.0f;
            // deferred chi2 (batch): this lane's experimental value and weight, fetched ahead of the row loop
            const bool with_terms = BATCH && eo.terms && (q0 + lane < nq);
            const float t_exp = with_terms ? M.expv[q0 + lane] : 0.0f;
            const float t_wts = (with_terms && M.wts) ? M.wts[q0 + lane] : 1.0f;
            if (resident && !(M.sq_exact & 2)) {
                // whole slab in shared memory, rows contiguous: flat loop over groups of 4 rows with an
                // 8-group (32-row) register prefetch, so the FADD chain never waits for a shared load
                if (!(M.sq_exact & 8)) mbar_wait(&mbar[0], 0u);
                const float4 *mt = reinterpret_cast<const float4 *>(ring) + lane;   // group g: mt[g*32]
                const float4 *gv = reinterpret_cast<const float4 *>(sG);            // group g: gv[g]
                const int ngroups = hs_pad / 4;                                      // multiple of 16
                // plain loads in a deeply unrolled loop: ptxas interleaves the LDS.128 of later groups with
                // the FADD chain of earlier ones (4.6 cycles/row measured, tools/ubench.cu)
#pragma unroll 1
                for (int g0 = 0; g0 < ngroups; g0 += 16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 a = mt[(g0 + j) * 32], g = gv[g0 + j];
                        acc = __fadd_rn(acc, __fmul_rn(g.x, a.x));      // rows beyond hs are zero rows: acc + 0
                        acc = __fadd_rn(acc, __fmul_rn(g.y, a.y));
                        acc = __fadd_rn(acc, __fmul_rn(g.z, a.z));
                        acc = __fadd_rn(acc, __fmul_rn(g.w, a.w));
                    }
                }
            } else
            for (int c = 0; c < ((M.sq_exact & 2) ? 0 : n_chunks); ++c) {
                const int stage = c % nst;
                mbar_wait(&mbar[stage], (unsigned)((c / nst) & 1));
                const float4 *mt = reinterpret_cast<const float4 *>(ring + stage * (SQ_ROWS * 32)) + lane;   // [r/4][lane]
                const float4 *gv = reinterpret_cast<const float4 *>(sG + c * SQ_ROWS);                     // [r/4]
                float prod[16], nxt[16];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 a = mt[u * 32], g = gv[u];
                    prod[4 * u] = __fmul_rn(g.x, a.x); prod[4 * u + 1] = __fmul_rn(g.y, a.y);
                    prod[4 * u + 2] = __fmul_rn(g.z, a.z); prod[4 * u + 3] = __fmul_rn(g.w, a.w);
                }
#pragma unroll
                for (int b = 0; b < SQ_ROWS / 4; b += 4) {
                    if (b + 4 < SQ_ROWS / 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 a = mt[(b + 4 + u) * 32], g = gv[b + 4 + u];
                            nxt[4 * u] = __fmul_rn(g.x, a.x); nxt[4 * u + 1] = __fmul_rn(g.y, a.y);
                            nxt[4 * u + 2] = __fmul_rn(g.z, a.z); nxt[4 * u + 3] = __fmul_rn(g.w, a.w);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, prod[u]);   // rows beyond hs are zero rows: acc + 0
                    if (b + 4 < SQ_ROWS / 4) {
#pragma unroll
                        for (int u = 0; u < 16; ++u) prod[u] = nxt[u];
                    }
                }
                if (c + nst < n_chunks) {              // slab larger than the ring: refill this stage
                    __syncwarp();                      // all lanes done with it
                    if (lane == 0) sq_issue(M, ring, mbar, slab, c + nst, n_chunks);
                }
            }
            if (q0 + lane < nq) {
                float sv = acc;
                if (M.kind == FRMC_KIND_SQ) {
                    sv = __fadd_rn(sv, 1.0f);
                    if (!(mrefit || M.prior || M.window) && mscale != 1.0f) sv = __fadd_rn(__fmul_rn(mscale, __fsub_rn(sv, 1.0f)), 1.0f);   // scale*(Sq-1)+1 (:775-778)
                } else {
                    if (!(mrefit || M.prior || M.window) && mscale != 1.0f) sv = __fmul_rn(mscale, sv);                            // (:1258-1260)
                }
                total[q0 + lane] = sv;
                if (with_terms) {
                    const float d = __fsub_rn(t_exp, sv);
                    float t = __fmul_rn(d, d);
                    if (M.wts) t = __fmul_rn(t_wts, t);
                    eo.terms[q0 + lane] = t;
                }
            }
            if (!(BATCH && eo.terms)) __threadfence();   // deferred: the grid barrier's fence follows at once
        }
        EPI_STAMP(3);
        if (BATCH && eo.terms) return;                 // uniform: every CTA sums the terms after the grid barrier
        // last CTA of this model (ticket) owns the chi^2
        __syncthreads();
        if (tid == 0) {
            const unsigned int t = atomicAdd(&tickets[m], 1u);
            s_last = (t == (unsigned int)(nblk - 1));
            if (s_last) tickets[m] = 0u;               // re-arm for the next launch
        }
        __syncthreads();
        EPI_STAMP(4);
        if (!s_last) return;
        __threadfence();
        if (mrefit || M.prior || M.window) {
            // ---- 3b. refit on S(Q)-1 against experimental-1 (StructureFactorConstraints.py:824-834, inherited by
            //          the reduced constraint), then scale the slices the other CTAs left unscaled, prior, window
            float sf = mscale;
            if (mrefit) {
                fit_scale_factor(es, sT, M, nq, [&](int i) { return __fsub_rn(__ldcg(total + i), 1.0f); },
                                 [&](int i) { return __fsub_rn(M.expv[i], 1.0f); });
                sf = es.s_sf;
            }
            for (int i = tid; i < nq; i += EPI_THREADS) {
                float sv = __ldcg(total + i);
                if (sf != 1.0f) sv = (M.kind == FRMC_KIND_SQ) ? __fadd_rn(__fmul_rn(sf, __fsub_rn(sv, 1.0f)), 1.0f) : __fmul_rn(sf, sv);
                if (M.prior) sv = __fadd_rn(M.prior[i], __fmul_rn(M.mf_weight, sv));
                total[i] = sv;
            }
            __threadfence();
            __syncthreads();
            if (M.window) {
                for (int i = tid; i < nq; i += EPI_THREADS)
                    sT[i] = convolve_same([&](int m) { return __ldcg(total + m); }, nq, M.window, M.n_window, i);
                __syncthreads();
                for (int i = tid; i < nq; i += EPI_THREADS) total[i] = sT[i];
                __threadfence();
                __syncthreads();
            }
        }
        for (int i = tid; i < nq; i += EPI_THREADS) {
            float d = __fsub_rn(M.expv[i], __ldcg(total + i));
            float t = __fmul_rn(d, d);
            if (M.wts) t = __fmul_rn(M.wts[i], t);
            sT[i] = t;
        }
        __syncthreads();
        chi2 = block_pairwise_sum(sT, M.pw_leaves, ps);
    }
    // ---- 4. publish
    if (stamps && tid == 0) {
        stamps[m * 8 + 5] = clock64();
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long *>(stamps + 122), gt);
    }
    if (BATCH) {
        // the engine's total standard error is formed by the decision walk of the batch kernel
        if (tid == 0) { eo.res[m] = chi2; eo.res[FRMC_MAX_MODELS + m] = es.s_sf; }
        return;
    }
    if (tid == 0) {
        chi2_out[m] = chi2;
        chi2_out[2 * FRMC_MAX_MODELS + m] = es.s_sf;       // the scale factor this evaluation used
        __threadfence_system();
        const unsigned int v = dev_seq[m] + 1u;
        dev_seq[m] = v;
        host_seq[m] = v;
    }
}

__global__ void __launch_bounds__(EPI_THREADS, 1)
epilogue_kernel(const ModelSet ms, GridSet gs, float *__restrict__ chi2_out,
                unsigned int *__restrict__ dev_seq, volatile unsigned int *__restrict__ host_seq,
                unsigned int *__restrict__ tickets, long long *__restrict__ stamps)
{
    extern __shared__ __align__(128) float epi_smem[];   // SQ: ring [n_stages][SQ_ROWS][32] | [hs_pad] G(r) | [out_pad] chi2 terms
    __shared__ __align__(16) EpiShared es;
    const int m = blockIdx.y, slab = blockIdx.x;
    if (m >= ms.n) return;
    const ModelDev &M = ms.m[m];
    const bool is_sq = (M.kind == FRMC_KIND_SQ || M.kind == FRMC_KIND_RSQ);
    if (slab >= (is_sq ? (M.n_out + 31) / 32 : 1)) return;
    EPI_STAMP(0);
    epilogue_prefetch(es, epi_smem, M, slab);
    epilogue_run<false>(es, epi_smem, ms, gs, m, slab, chi2_out, dev_seq, host_seq, tickets, stamps, EpiOut{});
}

// ------------------------------------------------------------------ kernels: fused Metropolis step
// ONE cooperative launch per proposal: (0) the epilogue CTAs start the TMA stream of their
// matrix slab, (1) every CTA helps resolve the PREVIOUS proposal (commit or clear the staged
// deltas, move the accepted atoms), grid barrier, (2) every CTA runs its share of the delta
// pass, grid barrier, (3) the first n_epi CTAs run the epilogue of their (model, Q slab), the
// others have exited.  Kernel boundaries cost ~3 us each on this part (profiles/), the two
// software grid barriers ~1 us each.
struct EpiMap {
    int n;                           // number of epilogue CTAs
    unsigned char model[128];
    unsigned char slab[128];
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// every thread of the CTA calls it after its last global write / atomic of the phase
__device__ __forceinline__ void grid_arrive(unsigned long long *bar)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(bar, 1ull);
}
__device__ __forceinline__ void grid_wait(const unsigned long long *bar, unsigned long long target)
{
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_u64(bar) < target)
            if (++spins > (1u << 28)) __trap();        // co-residency is guaranteed by the cooperative launch; never hang
    }
    __syncthreads();
}

__device__ __forceinline__ void resolve_body(const GridSet &gs, float4 *__restrict__ atoms, const Proposal *__restrict__ prop,
                                             const TotalsCopy &tc, int prev)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool accept = prev == 1;
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &G = gs.grid[gi];
        for (long long c = tid; c < 2 * G.cells; c += stride) {
            // __ldcg throughout: in the persistent kernel these cells were changed by other CTAs since this SM last
            // read them, and L1 is only coherent across kernel boundaries
            const int d = __ldcg(G.delta + c);
            if (d) { if (accept) G.counts[c] = (unsigned long long)((long long)__ldcg(G.counts + c) + d); G.delta[c] = 0; }
        }
        const long long ns = (long long)G.nsym * G.g.hs;
        if (accept) { for (long long c = tid; c < ns; c += stride) G.tot[c] = __ldcg(G.stot + c); }
        else        { for (long long c = tid; c < ns; c += stride) G.stot[c] = __ldcg(G.tot + c); }
    }
    if (accept) {
        if (tid < __ldcg(&prop->k)) atoms[__ldcg(&prop->pos[tid])] = __ldcg(&prop->newc[tid]);
        for (int m = 0; m < tc.n; ++m)
            for (long long i = tid; i < tc.len[m]; i += stride) tc.dst[m][i] = __ldcg(tc.src[m] + i);
    }
}

template <int MODE>
__global__ void __launch_bounds__(EPI_THREADS, 1)
propose_kernel(float4 *__restrict__ atoms, int npad, const ProposalIn in, Proposal *__restrict__ prop, Lattice L, GridSet gs,
               int nEl, unsigned long long *__restrict__ overflow, const ModelSet ms, const EpiMap em, int prev,
               const TotalsCopy tc, unsigned long long *__restrict__ bars, unsigned long long launch_no, unsigned long long resolve_no,
               float *__restrict__ chi2_out, unsigned int *__restrict__ dev_seq, volatile unsigned int *__restrict__ host_seq,
               unsigned int *__restrict__ tickets, long long *__restrict__ stamps)
{
    extern __shared__ __align__(128) float epi_smem[];
    __shared__ __align__(16) EpiShared es;
    __shared__ DeltaShared dsh;
    const bool epi = (int)blockIdx.x < em.n;
    const int m = epi ? em.model[blockIdx.x] : 0, slab = epi ? em.slab[blockIdx.x] : 0;
    if (stamps && threadIdx.x == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMin(reinterpret_cast<unsigned long long *>(stamps + 120), gt);
    }
    // (0) matrix slab stream
    if (epi) epilogue_prefetch(es, epi_smem, ms.m[m], slab);
    // barrier counters are monotonic across launches: bars[1] advances on every launch (launch_no),
    // bars[0] only on launches that resolve a previous proposal (resolve_no, counted by the host)
    const unsigned long long target = launch_no * gridDim.x;
    // (1) resolve the previous proposal
    if (prev) {
        resolve_body(gs, atoms, prop, tc, prev);
        grid_arrive(bars + 0);
        grid_wait(bars + 0, resolve_no * gridDim.x);
    }
    // (2) delta pass of this proposal
    delta_body<MODE>(dsh, atoms, npad, in, prop, L, gs, nEl, overflow);
    grid_arrive(bars + 1);
    if (stamps && threadIdx.x == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long *>(stamps + 121), gt);
    }
    if (!epi) return;
    grid_wait(bars + 1, target);
    // (3) epilogue
    EPI_STAMP(0);
    epilogue_run<false>(es, epi_smem, ms, gs, m, slab, chi2_out, dev_seq, host_seq, tickets, stamps, EpiOut{});
}

// ------------------------------------------------------------------ persistent variant of the per-move kernel
// One cooperative launch serves a whole run of proposals: the host writes a command (moved atoms + how the
// previous proposal was resolved) into mapped pinned memory, thread 0 of CTA 0 polls it, relays it through
// device memory to the other CTAs, and every CTA runs resolve -> delta pass -> (epilogue CTAs) G(r)/S(Q)/chi^2
// exactly like propose_kernel.  What disappears is the launch itself: ~4.5 us of cudaLaunchCooperativeKernel on
// the host and ~3 us of start-up on the device per evaluated move.  The kernel gives the GPU back when no
// command arrives for idle_ns (a watchdog, so a stalled or crashed host never leaves it spinning) or on QUIT.
enum : int { CMD_EVAL = 1, CMD_QUIT = 2 };

// Mapped pinned host memory.  A single-atom command fits two 16-byte chunks, each written by the host with ONE
// aligned 16-byte store and each carrying the command number, so the device validates them independently and a
// k = 1 proposal costs one PCIe read round trip to fetch; further atoms (k > 1) are read from pos/moved afterwards
// (the host fills those before the chunks).
struct alignas(64) HostCmd {
    unsigned int a_seq, a_word, a_pos0, a_mx0;   // chunk A: number, op | prev << 8 | k << 16, atom 0: position, moved x
    unsigned int b_my0, b_mz0, b_seq, b_pad;     // chunk B: atom 0 moved y, z, number again
    unsigned int alive;                          // written by the device: 1 while the kernel polls, 0 once it has left
    unsigned int pad[7];
    int pos[FRMC_MAX_GROUP];
    float moved[3 * FRMC_MAX_GROUP];
};

struct alignas(64) DevCmd {           // device memory: CTA 0's relay to the grid, same two self-validating chunks
    uint4 a, b;
    ProposalIn in;                    // atoms 1.. of a group move (written, and fenced, before the chunks)
};

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const unsigned int *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_v4(uint4 *p, uint4 v)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(EPI_THREADS, 1)
propose_loop_kernel(float4 *__restrict__ atoms, int npad, HostCmd *hcmd, DevCmd *dcmd, Proposal *__restrict__ prop, Lattice L,
                    GridSet gs, int nEl, unsigned long long *__restrict__ overflow, const ModelSet ms, const EpiMap em,
                    const TotalsCopy tc, unsigned long long *__restrict__ bars, unsigned int first_seq, unsigned long long idle_ns,
                    float *__restrict__ chi2_out, unsigned int *__restrict__ dev_seq, volatile unsigned int *__restrict__ host_seq,
                    unsigned int *__restrict__ tickets)
{
    extern __shared__ __align__(128) float epi_smem[];
    __shared__ __align__(16) EpiShared es;
    __shared__ DeltaShared dsh;
    __shared__ ProposalIn s_in;
    __shared__ uint4 s_a, s_b;             // the command's two chunks
    const bool epi = (int)blockIdx.x < em.n;
    const int m = epi ? em.model[blockIdx.x] : 0, slab = epi ? em.slab[blockIdx.x] : 0;
    if (epi) epilogue_prefetch(es, epi_smem, ms.m[m], slab);      // the slab stays resident for the whole run
    unsigned int expect = first_seq;
    unsigned long long n_delta = 0, n_resolve = 0;                // barrier generations, counted identically by every CTA
    for (;;) {
        // (a) CTA 0 fetches the next command from the host and relays it
        if (blockIdx.x == 0) {
            if (threadIdx.x == 0) {
                unsigned long long t0, t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                uint4 A = make_uint4(expect, (unsigned)CMD_QUIT, 0u, 0u), B = make_uint4(0u, 0u, expect, 0u);
                for (;;) {
                    const uint4 hA = ld_volatile_v4(&hcmd->a_seq), hB = ld_volatile_v4(&hcmd->b_my0);
                    if (hA.x == expect && hB.z == expect) { A = hA; B = hB; break; }
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > idle_ns) break;           // idle: give the GPU back (relayed as QUIT)
                }
                s_a = A; s_b = B;
            }
            __syncthreads();
            const int k = (int)(s_a.y >> 16);
            if ((int)(s_a.y & 0xFFu) == CMD_EVAL && k > 1) {
                __threadfence_system();                     // entries 1.. were written before the chunks
                for (int t = 1 + threadIdx.x; t < k; t += blockDim.x) {
                    dcmd->in.pos[t] = (int)ld_volatile_u32((const unsigned int *)&hcmd->pos[t]);
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        dcmd->in.moved[3 * t + c] = __uint_as_float(ld_volatile_u32((const unsigned int *)&hcmd->moved[3 * t + c]));
                }
                __threadfence();
                __syncthreads();
            }
            if (threadIdx.x == 0) { st_volatile_v4(&dcmd->b, s_b); st_volatile_v4(&dcmd->a, s_a); }
        }
        // (b) every CTA picks the command up: one L2 round trip for a single-atom move
        if (threadIdx.x == 0) {
            uint4 A, B;
            do { A = ld_volatile_v4(&dcmd->a.x); B = ld_volatile_v4(&dcmd->b.x); } while (A.x != expect || B.z != expect);
            s_a = A; s_b = B;
        }
        __syncthreads();
        const int op = (int)(s_a.y & 0xFFu), prev = (int)((s_a.y >> 8) & 0xFFu);
        if (op != CMD_EVAL) break;
        {
            const int k = (int)(s_a.y >> 16);
            if (threadIdx.x == 0) {
                s_in.k = k; s_in.pos[0] = (int)s_a.z;
                s_in.moved[0] = __uint_as_float(s_a.w); s_in.moved[1] = __uint_as_float(s_b.x); s_in.moved[2] = __uint_as_float(s_b.y);
            }
            if (k > 1) {
                __threadfence();
                for (int t = 1 + threadIdx.x; t < k; t += blockDim.x) {
                    s_in.pos[t] = __ldcg(&dcmd->in.pos[t]);
#pragma unroll
                    for (int c = 0; c < 3; ++c) s_in.moved[3 * t + c] = __ldcg(&dcmd->in.moved[3 * t + c]);
                }
            }
        }
        __syncthreads();
        // (1) resolve the previous proposal
        if (prev) {
            resolve_body(gs, atoms, prop, tc, prev);
            grid_arrive(bars + 0);
            ++n_resolve;
            grid_wait(bars + 0, n_resolve * gridDim.x);
        }
        // (2) delta pass of this proposal
        delta_body<MODE>(dsh, atoms, npad, s_in, prop, L, gs, nEl, overflow);
        grid_arrive(bars + 1);
        ++n_delta;
        // (3) epilogue; the other CTAs go straight back to waiting (the host sends the next command only after chi^2)
        if (epi) {
            grid_wait(bars + 1, n_delta * gridDim.x);
            epilogue_run<false>(es, epi_smem, ms, gs, m, slab, chi2_out, dev_seq, host_seq, tickets, nullptr, EpiOut{});
        }
        ++expect;
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *reinterpret_cast<volatile unsigned int *>(&hcmd->alive) = 0u;
        __threadfence_system();
    }
}


// ------------------------------------------------------------------ a run of proposals resolved on the device
// Engine.__on_runtime_step_try_move (Engine.py:3302-3338) for a whole RUN of proposals in one cooperative launch:
// the proposals' before/after deltas come from ONE pass over the store (every record is tested against the old and
// the new position of every moved atom of every proposal: 16 B/atom of traffic for the whole batch), and the
// sequential part of the Metropolis chain -- chi^2 of proposal j depends on which earlier proposals were accepted --
// is resolved on the device in ROUNDS:
//   * G groups of epilogue CTAs evaluate the next G unresolved proposals in parallel, each on
//     (committed totals + its own delta);
//   * after a grid barrier every CTA walks their total standard errors in order with the engine's rule
//     (new > old: rejected unless the next pre-drawn random number <= tolerance) up to the first accepted one;
//     evaluations made beyond it are void and are repeated in the next round;
//   * the accepted proposal is committed by all CTAs (ordered counts, symmetrised totals, atoms, model totals) and
//     the deltas of the proposals behind it are corrected for the pairs (their atoms) x (the atoms that just
//     moved): those pairs were formed against the positions at the start of the batch.
// A proposal that moves an atom already moved by an accepted proposal of the same batch cannot be corrected; the
// batch stops in front of it (BatchRun::stopped) and the host starts the next batch there.
// The result is the sequential path's, bit for bit: the counts are integers and each evaluation sees exactly the
// totals the sequential path would have staged.
static const int BATCH_MAX_PROPS = 32;
static const int BATCH_MAX_GROUPS = 16;
static const int BATCH_MAX_SPEC = 5;
static const int BATCH_DEFER_MAX_LEAVES = 32;   // longest pairwise schedule a warp sums (n_out up to ~4000)   // accepted-but-uncommitted proposals an evaluation may assume (EpiOut::add2)
static const int BATCH_PER_LAUNCH = 128;   // batches of host proposals one launch works through (~20 ms)
static const int BATCH_STAMP_SLOTS = 4 + 5 * 64;   // start, cleared, delta pass done, end; 5 per round
static const int BATCH_STAMP_TOTAL = BATCH_STAMP_SLOTS + 64 * 128;   // + per round the EPI_STAMP block of CTA 1 (see tools/probe_batch.py)

struct BatchIn {                      // by value
    int n_prop, n_atoms;
    int out_base;                     // index of proposal 0 in the call-wide output arrays
    int pad;
    int first[BATCH_MAX_PROPS + 1];   // atoms of proposal j: [first[j], first[j+1])
    unsigned int share[BATCH_MAX_PROPS];   // bit i: earlier proposal i moves an atom of proposal j
    int pos[FRMC_MAX_GROUP];          // positions in the sorted store
    float moved[3 * FRMC_MAX_GROUP];
};

struct BatchRun {                     // device memory, carried from launch to launch of one call
    float total;                      // the engine's totalStandardError
    int n_rand;                       // random numbers consumed
    int stopped;                      // a conflict ended the run at proposal n_done
    int n_done;                       // proposals resolved (call-wide index)
    int n_accepted;
    int rounds;                       // evaluation rounds (statistics)
    int gen_state;                    // device-generated runs: 0 running, 1 every proposal of the call is resolved, 3 a generated
                                      // coordinate left the window the geometry mode was chosen for (the host re-plans)
    int pad;
    float cchi2[FRMC_MAX_MODELS];     // chi2 per model of the committed state
    float csf[FRMC_MAX_MODELS];
};

// Pair corrections applied ON THE FLY (launches of single-atom proposals).  The delta of proposal x was formed against
// the positions at the start of the launch; if an earlier proposal y of the launch is accepted and their atoms are in
// range, x's true delta differs by the four events +h(x_old,y_old) -h(x_old,y_new) -h(x_new,y_old) +h(x_new,y_new).
// The head writes them per (x, y, grid) into a table; a node that ASSUMES y accepted adds them to the counts it loads
// (so a chain of predicted acceptances may run through atoms that see each other: in a small dense system, where every
// pair of atoms is in range, that is the difference between one and five proposals per round), and the commit of
// several acceptances of one round adds them to the totals.  Acceptances already committed are handled as before: the
// deltas of the proposals behind them are corrected in place.
struct CorrEv {
    int sym;                          // (symmetrised row << 16) | bin, or -1: the event changes no histogram cell
    int ord;                          // ((ordered cell index) << 3) | overflow code << 1 | (sign > 0); overflow code: 0 none,
};                                    // 1 counts +1, 2 counts -1 towards the proposal's edge-overflow events
static const int CORR_PER_PAIR = FRMC_MAX_GRIDS * 4;
static const int CORR_NODE_MAX = 64;              // events one node applies (<= 15 pairs x 4 on its model's grid)
static const int CORR_COMMIT_MAX = 240;           // events one commit applies (<= 15 pairs x 4 x grids)

struct BatchDev {                     // by value: the batch's device buffers
    int *bsym[FRMC_MAX_GRIDS];        // [BATCH_MAX_PROPS][nsym*hs]  symmetrised delta per proposal
    int *bdelta[FRMC_MAX_GRIDS];      // [BATCH_MAX_PROPS][2*cells]  ordered delta per proposal
    float *btotal[FRMC_MAX_MODELS];   // [BATCH_MAX_PROPS][n_out]   model totals per evaluation slot (two alternating sets of BATCH_MAX_GROUPS)
    float *bterm[FRMC_MAX_MODELS];    // [BATCH_MAX_PROPS][n_out]   chi2 terms per evaluation slot of the models in defer_mask
    unsigned int defer_mask;          // S(Q) models whose chi2 every CTA sums after the grid barrier (EpiOut::terms)
    float *total_committed[FRMC_MAX_MODELS];
    float *res;                       // [BATCH_MAX_PROPS][2*FRMC_MAX_MODELS]
    float *ptotal;                    // [BATCH_MAX_PROPS]
    unsigned int *tickets;            // [BATCH_MAX_PROPS][FRMC_MAX_MODELS] slab tickets, then [BATCH_MAX_PROPS] model tickets
    unsigned long long *bov;          // [BATCH_MAX_PROPS] edge-overflow events per proposal
    BatchRun *run;
    const float *rand;                // call-wide
    float *out_chi2;                  // call-wide [n][n_models]
    int *out_dec;                     // call-wide [n]: 0 rejected, 1 accepted, 2 accepted within the tolerance
    float var2[FRMC_MAX_MODELS];
    float tol;
    float box_eps;                    // rounding margin of a sub-block box: 1e-6 (1 + largest |coordinate| of the store and the proposals)
    int n_groups;                     // G
    int freq[FRMC_MAX_MODELS];        // scale-factor refit schedule per model (0: none): an evaluation refits when the engine's
    unsigned long long accepted_base; // accepted count at its node (accepted_base + acceptances of this call so far) % freq == 0
    CorrEv *corr;                     // [BATCH_MAX_PROPS * (BATCH_MAX_PROPS - 1) / 2][FRMC_MAX_GRIDS][4], pair (x, y < x) at x (x - 1) / 2 + y
    int debug;                        // FRMC_BATCH_DEBUG: timing experiments only (results are wrong when set)
    int rand_per_proposal;            // 1: rand[i] belongs to proposal i of the call (counter-based contract, fullrmc_b200/rng.py);
                                      // 0: consumed in order, one per worse proposal (the reference's generate_random_float stream)
};

// ---- device-generated runs: a launch's proposals are drawn on the device
struct GenOut {                       // device memory: what the generator leaves for the commit besides the BatchIn
    int ridx[FRMC_MAX_GROUP];         // real index of every moved atom
    float mreal[3 * FRMC_MAX_GROUP];  // moved REAL coordinates (engine.realCoordinates of an accepted move, Engine.py:3337)
};

struct GenParams {                    // by value
    unsigned long long seed, first_counter;
    int n_total;                      // proposals of the call
    int n_groups;
    const int *goff, *gidx;           // groups: atoms gidx[goff[g] .. goff[g+1]) (real indexes)
    const int *inv;                   // real index -> position in the sorted store
    const float4 *real;               // real coordinates by real index (periodic systems)
    float rb[9];                      // reciprocalBasisVectors (transform_coordinates' transMatrix)
    int pbc;
    float amp_min, amp_max;
    float win_lo[3], win_hi[3];       // box coordinates must stay inside (the geometry mode and culling margin assume it)
    float *rand_out;                  // call-wide: acceptance number of every proposal
    int *group_out;                   // call-wide: selected group of every proposal (may be NULL)
};

struct BatchShared {
    float4 sOld[FRMC_MAX_GROUP];
    float4 sNew[FRMC_MAX_GROUP];
    int sPos[FRMC_MAX_GROUP];
    int sProp[FRMC_MAX_GROUP];
    float4 fOld[FRMC_MAX_GROUP];               // low / high corner of the box spanned by a moved atom's old and new position
    float4 fNew[FRMC_MAX_GROUP];               // (blocks_far's conventions: periodically reduced coordinates, lo.w = rounding margin)
    float s_pt[BATCH_MAX_GROUPS], s_rand[2 * BATCH_MAX_GROUPS];   // random numbers from the round's first: the walk's, then the plan's
    unsigned int ev[CORR_NODE_MAX];            // the node's on-the-fly pair corrections (EpiOut::ev)
    unsigned int ev_bits[32];                  // filter of the node's list, then of the commit's list
    int n_ev, n_cev;
    float s_csf[FRMC_MAX_MODELS];              // committed scale factor per model as the launch proceeds (refit schedules)
    unsigned int s_cnt_mod[FRMC_MAX_MODELS];   // (accepted count at the start of the launch) % refit frequency, per model
    int in_first[BATCH_MAX_PROPS + 1];         // the launch's BatchIn fields the rounds read (first, share): staged once, so the
    unsigned int in_share[BATCH_MAX_PROPS];    // plan and the walk never wait for a parameter / global load
    float s_prand[BATCH_MAX_PROPS];            // rand_per_proposal: the acceptance number of proposal j of this launch
    unsigned int near[BATCH_MAX_PROPS];        // bit i of near[j]: a pair (atom of j, atom of earlier proposal i) is in range
    // the round's plan (rebuilt after every walk by thread 0): slot s evaluates proposal slot_k[s] on the committed
    // state plus the proposals of slot_A[s]; child = slot of the next proposal after a rejection [0] / an acceptance
    // [1], or -1
    int slot_k[BATCH_MAX_GROUPS];
    unsigned int slot_A[BATCH_MAX_GROUPS];
    int slot_child[BATCH_MAX_GROUPS][2];
    int n_slots;
    int path_len;                              // slots 0 .. path_len-1 are linked one after the other ...
    unsigned int path_exp;                     // ... by these expected decisions (bit s: slot s accepted)
    float est[BATCH_MAX_PROPS];                // predicted change of the total standard error by proposal j (has_est bit j)
    unsigned int has_est;
    int sched[FRMC_MAX_MODELS][4 * BATCH_DEFER_MAX_LEAVES];   // pairwise-summation schedules of the models in defer_mask
    unsigned int symmask[BATCH_MAX_PROPS];     // rows of the symmetrised delta proposal j can touch (all ones when unknown)
    float s_chi[BATCH_MAX_GROUPS][FRMC_MAX_MODELS];                // chi2 per slot and model of the round
    float pw_scratch[EPI_THREADS / 32][BATCH_DEFER_MAX_LEAVES];    // leaf sums of the per-warp pairwise summation
    int path_slot[BATCH_MAX_GROUPS], path_dec[BATCH_MAX_GROUPS];   // the walk of the round: slot and decision per proposal
    int s_last;
    unsigned int s_acc;
    int s_cur, s_ri, s_stopped;
    float s_total;
    unsigned long long s_bar;
};

// block_pairwise_sum for ONE warp over v[0..n) in GLOBAL memory (written by other CTAs before a grid barrier); the
// schedule (leaf offsets | leaf lengths | combine dst | combine src) in shared memory.  Four leaves at a time, 8 lanes
// per leaf = the 8 accumulators of numpy's unrolled leaf loop (leaves are at most 128 long: 16 elements per lane).
// Every element a lane needs is loaded before the first addition (one L2 round trip per four leaves).
// All 32 lanes call; result valid in lane 0.
__device__ __forceinline__ float warp_pairwise_sum(const float *__restrict__ v, const int *sched, int nl, float *leaf_sum)
{
    const int lane = threadIdx.x & 31, group = lane >> 3, lane8 = lane & 7;
    for (int base = 0; base < nl; base += 4) {
        const int l = base + group;
        const bool live = l < nl;
        const int off = live ? sched[l] : 0, len = live ? sched[nl + l] : 0;
        const bool big = len >= 8;
        const int main_len = big ? len - (len % 8) : 0;
        float x[16], rem[7];
#pragma unroll
        for (int u = 0; u < 16; ++u) x[u] = (8 * u < main_len) ? __ldcg(v + off + 8 * u + lane8) : 0.0f;
#pragma unroll
        for (int u = 0; u < 7; ++u) rem[u] = (live && lane8 == 0 && main_len + u < len) ? __ldcg(v + off + main_len + u) : 0.0f;
        float r = x[0];
#pragma unroll
        for (int u = 1; u < 16; ++u) if (8 * u < main_len) r = __fadd_rn(r, x[u]);
        __syncwarp();
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 4));
        float res = big ? r : 0.0f;
        if (live && lane8 == 0) {
#pragma unroll
            for (int u = 0; u < 7; ++u) if (main_len + u < len) res = __fadd_rn(res, rem[u]);
            leaf_sum[l] = res;
        }
    }
    __syncwarp();
    float out = 0.f;
    if (lane == 0) {
        for (int i = 0; i < nl - 1; ++i) {
            const int d = sched[2 * nl + i], sr = sched[3 * nl + i];
            leaf_sum[d] = __fadd_rn(leaf_sum[d], leaf_sum[sr]);
        }
        out = leaf_sum[0];
    }
    __syncwarp();
    return out;
}

// one signed event of proposal `j` (the batch's delta_hit)
__device__ __forceinline__ void batch_hit(float d2, int sign, int same, int slab, int sym, int j, const GridSet &gs, const BatchDev &bd,
                                          int nEl, unsigned long long &ov)
{
#pragma unroll 1
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &G = gs.grid[gi];
        if (in_range(d2, G.g)) {
            int *bdel = bd.bdelta[gi] + (long long)j * 2 * G.cells;
            int *bsy = bd.bsym[gi] + (long long)j * G.nsym * G.g.hs;
            const int b = bin_index(d2, G.g);
            if (b < G.g.hs) {
                atomicAdd(&bdel[(same ? 0 : G.cells) + (long long)slab * G.g.hs + b], sign);
                atomicAdd(&bsy[(long long)sym * G.g.hs + b], sign);
            } else {
                ++ov;
                const long long flat = (long long)slab * G.g.hs + b;
                if (G.g.spill && flat < G.cells) {
                    const int s2 = (int)(flat / G.g.hs), b2 = (int)(flat - (long long)s2 * G.g.hs);
                    atomicAdd(&bdel[(same ? 0 : G.cells) + flat], sign);
                    atomicAdd(&bsy[(long long)sym_index(s2 / nEl, s2 % nEl, nEl) * G.g.hs + b2], sign);
                }
            }
        }
    }
}

// the histogram event of one distance on grid G as batch_hit would register it, as a CorrEv (sign folded in later)
__device__ __forceinline__ CorrEv corr_event(float d2, int same, int slab, int sym, const GridDev &G, int nEl)
{
    CorrEv ev;
    ev.sym = -1; ev.ord = 0;
    if (!in_range(d2, G.g)) return ev;
    const int b = bin_index(d2, G.g);
    if (b < G.g.hs) {
        ev.sym = (sym << 16) | b;
        ev.ord = (int)(((same ? 0 : G.cells) + (long long)slab * G.g.hs + b) << 3);
        return ev;
    }
    ev.ord = 2;                                   // an edge-overflow event (code 1; the caller flips it for "undone" pairs)
    const long long flat = (long long)slab * G.g.hs + b;
    if (G.g.spill && flat < G.cells) {            // the reference's unchecked write lands in the next slab
        const int s2 = (int)(flat / G.g.hs), b2 = (int)(flat - (long long)s2 * G.g.hs);
        ev.sym = (sym_index(s2 / nEl, s2 % nEl, nEl) << 16) | b2;
        ev.ord |= (int)(((same ? 0 : G.cells) + flat) << 3);
    }
    return ev;
}

// order-preserving map float -> unsigned (and back): min / max of floats through integer REDUX
__device__ __forceinline__ unsigned int float_order(float f)
{
    const unsigned int b = __float_as_uint(f);
    return b ^ ((b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float order_float(unsigned int k)
{
    return __uint_as_float(k ^ ((k & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu));
}

// a position as a one-point box for blocks_far (block_bbox_kernel's conventions: fractional part under PBC,
// w = margin for the rounding of fl(xi - xj) on the raw coordinates)
__device__ __forceinline__ float4 point_box(float x, float y, float z, int pbc)
{
    const float amax = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    if (pbc) { x -= floorf(x); y -= floorf(y); z -= floorf(z); }
    return make_float4(x, y, z, 1e-6f * (1.0f + amax));
}

__device__ __forceinline__ void zero_ints(int *p, long long n)
{
    // p is 16-byte aligned and the allocation ends on a multiple of four ints (the host pads it)
    int4 *q = reinterpret_cast<int4 *>(p);
    const long long n4 = (n + 3) >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        q[i] = make_int4(0, 0, 0, 0);
}

// One launch's proposals drawn on the device (one warp; lane j = proposal n_done + j of the call): group from the
// step's first word, translation vector, moved real coordinates = current real coordinates + vector, moved box
// coordinates = transform_coordinates(reciprocal basis, moved real) -- Engine.__on_runtime_step_select_group,
// Engine.py:3166-3228, for translation generators -- with the counter-based numbers of rng.cuh.  The current
// coordinates are read AFTER every earlier launch has committed (stream order), and a proposal that moves an atom an
// accepted proposal of its own launch moved ends that launch (batch_kernel), so it is generated again from the
// new position: the run equals the sequential one.
__global__ void generate_batch_kernel(const GenParams gp, BatchRun *__restrict__ run, BatchIn *__restrict__ din, GenOut *__restrict__ gen,
                                      const float4 *__restrict__ atoms)
{
    __shared__ int s_k[BATCH_MAX_PROPS], s_first[BATCH_MAX_PROPS + 1], s_g[BATCH_MAX_PROPS];
    __shared__ int s_pos[FRMC_MAX_GROUP];
    __shared__ int s_np, s_bad;
    const int j = threadIdx.x;
    const int base = run->n_done;
    if (j == 0) s_bad = 0;
    if (base >= gp.n_total || run->gen_state == 3) {
        if (j == 0) { din->n_prop = 0; din->n_atoms = 0; din->out_base = base; if (base >= gp.n_total) run->gen_state = 1; }
        return;
    }
    const bool active = base + j < gp.n_total;
    const StepRandom sr = step_random(gp.seed, gp.first_counter + (unsigned long long)(base + j), gp.amp_min, gp.amp_max);
    const int g = (int)(((unsigned long long)sr.group_word * (unsigned long long)gp.n_groups) >> 32);
    const int k = active ? gp.goff[g + 1] - gp.goff[g] : 0;
    s_k[j] = k; s_g[j] = g;
    __syncwarp();
    if (j == 0) {
        int na = 0, np = 0;
        while (np < BATCH_MAX_PROPS && s_k[np] > 0 && na + s_k[np] <= FRMC_MAX_GROUP) { s_first[np] = na; na += s_k[np]; ++np; }
        s_first[np] = na;
        s_np = np;
        din->n_prop = np; din->n_atoms = na; din->out_base = base;
        for (int i = 0; i <= np; ++i) din->first[i] = s_first[i];
    }
    __syncwarp();
    const int np = s_np;
    if (j < np) {
        bool bad = false;
        for (int t = 0; t < k; ++t) {
            const int r = gp.gidx[gp.goff[g] + t];
            const int pos = gp.inv[r];
            float x, y, z;
            if (gp.pbc) { const float4 rc = gp.real[r]; x = rc.x; y = rc.y; z = rc.z; }
            else { const float4 rc = atoms[pos]; x = rc.x; y = rc.y; z = rc.z; }      // non-periodic: box coordinates ARE the real ones
            const float mx = __fadd_rn(x, sr.vx), my = __fadd_rn(y, sr.vy), mz = __fadd_rn(z, sr.vz);
            float bx = mx, by = my, bz = mz;
            if (gp.pbc) transform_point(gp.rb, mx, my, mz, bx, by, bz);
            const int a = s_first[j] + t;
            s_pos[a] = pos;
            din->pos[a] = pos;
            din->moved[3 * a] = bx; din->moved[3 * a + 1] = by; din->moved[3 * a + 2] = bz;
            gen->ridx[a] = r;
            gen->mreal[3 * a] = mx; gen->mreal[3 * a + 1] = my; gen->mreal[3 * a + 2] = mz;
            bad = bad || !(bx >= gp.win_lo[0] && bx <= gp.win_hi[0] && by >= gp.win_lo[1] && by <= gp.win_hi[1] &&
                           bz >= gp.win_lo[2] && bz <= gp.win_hi[2]);
        }
        if (bad) atomicExch(&s_bad, 1);
        gp.rand_out[base + j] = sr.accept;
        if (gp.group_out) gp.group_out[base + j] = g;
    }
    __syncwarp();
    if (j < np) {
        unsigned int share = 0u;                 // earlier proposals of the launch that move one of this proposal's atoms
        for (int t = s_first[j]; t < s_first[j + 1]; ++t)
            for (int v = 0; v < s_first[j]; ++v)
                if (s_pos[v] == s_pos[t]) {
                    int jp = 0;
                    while (s_first[jp + 1] <= v) ++jp;
                    share |= 1u << jp;
                }
        din->share[j] = share;
    }
    __syncwarp();
    if (j == 0) {
        if (s_bad) { din->n_prop = 0; din->n_atoms = 0; run->gen_state = 3; }
        run->stopped = 0;
    }
}

template <int MODE, bool GEN, bool FLYT>
__global__ void __launch_bounds__(EPI_THREADS, 1)
batch_kernel(float4 *__restrict__ atoms, int npad, const BatchIn *__restrict__ in_arr, int n_batches, Lattice L, GridSet gs, int nEl, const ModelSet ms,
             const EpiMap em, const BatchDev bd, const CullParams cp, unsigned long long *__restrict__ bars,
             unsigned long long *__restrict__ overflow, long long *__restrict__ stamps, const BatchIn *__restrict__ in_dev,
             const GenOut *__restrict__ gen, float4 *__restrict__ real)
{
    extern __shared__ __align__(128) float epi_smem[];
    __shared__ __align__(16) EpiShared es;
    __shared__ BatchShared bs;
    const int tid = threadIdx.x;
    // The proposals: host proposals arrive as an ARRAY of batches in device memory (in_arr[0 .. n_batches)), and ONE launch
    // works through all of them -- a launch per batch cost 16 us of launch latency, prologue and S(Q)-slab staging around
    // 148 us of work; generated proposals are drawn batch by batch by generate_batch_kernel (in_dev, one batch per launch;
    // gen / real then carry the real coordinates to commit).
    const int nb = GEN ? 1 : n_batches;
    if (nb <= 0) return;
    if (GEN && in_dev->n_prop == 0) return;          // nothing generated: the call is finished or must be re-planned (uniform)
    if (__ldcg(&bd.run->stopped)) return;            // an earlier launch of this call hit a conflict: nothing to do (uniform)
    // debug timeline (FRMC_BATCH_STAMPS=1): globaltimer ns of CTA 0 at the phase boundaries of the LAST batch
#define BATCH_STAMP(i) do { if (stamps && blockIdx.x == 0 && tid == 0 && (i) < BATCH_STAMP_SLOTS) { \
        unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); stamps[(i)] = (long long)gt_; } } while (0)
    const int G = bd.n_groups;
    const int group = (int)blockIdx.x / em.n, e_idx = (int)blockIdx.x - group * em.n;
    const bool epi = group < G;
    const int m = epi ? em.model[e_idx] : 0, slab = epi ? em.slab[e_idx] : 0;
    if (epi) epilogue_prefetch(es, epi_smem, ms.m[m], slab);
    // grid barrier generations: the counter only grows; every earlier launch left it at a multiple of the grid size
    // and fewer than gridDim.x CTAs can have arrived at this launch's first barrier when this one reads it
    if (tid == 0) {
        const unsigned long long v = ld_acquire_u64(bars);
        bs.s_bar = v - v % gridDim.x;
    }
    __syncthreads();
    unsigned long long bar_target = bs.s_bar;
  for (int batch = 0; batch < nb; ++batch) {
    const BatchIn &in = GEN ? *in_dev : in_arr[batch];
    if (batch > 0) {
        // the previous batch is behind every CTA (its commits, the run state CTA 0 wrote); a conflict ends the launch
        grid_arrive(bars);
        bar_target += gridDim.x;
        grid_wait(bars, bar_target);
        if (__ldcg(&bd.run->stopped)) return;        // uniform: written before the barrier
    }
    BATCH_STAMP(0);
    const int np = in.n_prop, na = in.n_atoms;
    // every proposal moves one atom: pair corrections between proposals of the launch are applied on the fly (CorrEv)
    // (FLYT is a template parameter: the host leaves the machinery out for large sparse systems, where two proposals of
    // a launch hardly ever see each other and the leaner kernel is a few percent faster)
    const bool fly = FLYT && (na == np) && np > 1 && bd.corr != nullptr;
    // ---- (1) clear the per-proposal buffers, load the moved atoms
    for (int gi = 0; gi < gs.n; ++gi) {
        const GridDev &Gd = gs.grid[gi];
        zero_ints(bd.bsym[gi], (long long)np * Gd.nsym * Gd.g.hs);
        zero_ints(bd.bdelta[gi], (long long)np * 2 * Gd.cells);
    }
    if (blockIdx.x == 0) {
        for (int i = tid; i < BATCH_MAX_PROPS * (FRMC_MAX_MODELS + 1); i += blockDim.x) bd.tickets[i] = 0u;
        for (int i = tid; i < BATCH_MAX_PROPS; i += blockDim.x) bd.bov[i] = 0ull;
    }
    for (int t = tid; t < na; t += blockDim.x) {
        const int p = in.pos[t];
        const float4 o = __ldcg(atoms + p);
        bs.sOld[t] = o;
        bs.sNew[t] = make_float4(in.moved[3 * t], in.moved[3 * t + 1], in.moved[3 * t + 2], o.w);
        bs.sPos[t] = p;
        {
            const float4 fo = point_box(o.x, o.y, o.z, cp.pbc), fn = point_box(in.moved[3 * t], in.moved[3 * t + 1], in.moved[3 * t + 2], cp.pbc);
            bs.fOld[t] = make_float4(fminf(fo.x, fn.x), fminf(fo.y, fn.y), fminf(fo.z, fn.z), fmaxf(fo.w, fn.w));
            bs.fNew[t] = make_float4(fmaxf(fo.x, fn.x), fmaxf(fo.y, fn.y), fmaxf(fo.z, fn.z), 0.f);
        }
        int j = 0;
        while (j + 1 < np && in.first[j + 1] <= t) ++j;
        bs.sProp[t] = j;
    }
    if (tid < BATCH_MAX_PROPS) {
        bs.near[tid] = 0u; bs.symmask[tid] = 0u;
        bs.s_prand[tid] = (bd.rand_per_proposal && tid < np) ? __ldcg(bd.rand + in.out_base + tid) : 0.0f;
        if (tid < FRMC_MAX_MODELS) {
            bs.s_csf[tid] = (tid < ms.n) ? __ldcg(&bd.run->csf[tid]) : 1.0f;
            bs.s_cnt_mod[tid] = (tid < ms.n && bd.freq[tid] > 0)
                ? (unsigned int)((bd.accepted_base + (unsigned long long)__ldcg(&bd.run->n_accepted)) % (unsigned long long)bd.freq[tid]) : 0u;
        }
        bs.in_share[tid] = (tid < np) ? in.share[tid] : 0u;
        bs.in_first[tid] = in.first[min(tid, np)];
        if (tid == 0) bs.in_first[BATCH_MAX_PROPS] = in.first[min(BATCH_MAX_PROPS, np)];
    }
    for (int mm = 0; mm < ms.n; ++mm)
        if ((bd.defer_mask >> mm) & 1u)
            for (int i = tid; i < 4 * ms.m[mm].pw_leaves; i += blockDim.x) bs.sched[mm][i] = ms.m[mm].pw_sched[i];
    __syncthreads();
    {
        // rows [el_t, *] of the symmetrised delta are the only ones proposal j can touch -- unless there are more than
        // 32 rows or a grid reproduces the reference's edge-bin spill (the event lands in the next slab)
        bool any = nEl * (nEl + 1) / 2 > 32;
        for (int gi = 0; gi < gs.n; ++gi) any = any || gs.grid[gi].g.spill;
        for (int t = tid; t < na; t += blockDim.x) {
            unsigned int mk = 0xFFFFFFFFu;
            if (!any) {
                mk = 0u;
                const int et = (int)(__float_as_uint(bs.sOld[t].w) & 0xFF);
                for (int e2 = 0; e2 < nEl; ++e2) mk |= 1u << sym_index(et, e2, nEl);
            }
            atomicOr(&bs.symmask[bs.sProp[t]], mk);
        }
    }
    // which earlier proposals would change proposal j's delta if they were accepted (the pairs the commit corrects)
    for (int e = tid; e < na * na; e += blockDim.x) {
        const int t = e / na, u = e - t * na;
        const int jt = bs.sProp[t], ju = bs.sProp[u];
        if (ju >= jt) continue;
        const float4 ot = bs.sOld[t], nt = bs.sNew[t], ou = bs.sOld[u], nu = bs.sNew[u];
        const float d_oo = dist2<MODE>(ot.x, ot.y, ot.z, ou.x, ou.y, ou.z, L);
        const float d_on = dist2<MODE>(ot.x, ot.y, ot.z, nu.x, nu.y, nu.z, L);
        const float d_no = dist2<MODE>(nt.x, nt.y, nt.z, ou.x, ou.y, ou.z, L);
        const float d_nn = dist2<MODE>(nt.x, nt.y, nt.z, nu.x, nu.y, nu.z, L);
        const bool hit = ((d_oo >= gs.t2lo) && (d_oo < gs.t2hi)) || ((d_on >= gs.t2lo) && (d_on < gs.t2hi)) ||
                         ((d_no >= gs.t2lo) && (d_no < gs.t2hi)) || ((d_nn >= gs.t2lo) && (d_nn < gs.t2hi));
        if (hit) atomicOr(&bs.near[jt], 1u << ju);
        if (hit && fly && blockIdx.x == 0) {
            // the four events of the pair per grid (x = the later proposal's atom, y = the earlier one's): +oo -on -no +nn;
            // overflow events count towards x's total with +1 for the re-done pairs (on, nn) and -1 for the undone (oo, no)
            const uint32_t mt = __float_as_uint(ot.w), mu = __float_as_uint(ou.w);
            const int same = (mt >> 8) == (mu >> 8);
            const int et = (int)(mt & 0xFF), eu = (int)(mu & 0xFF);
            const int slab_ = et * nEl + eu, sym = sym_index(et, eu, nEl);
            const float d4[4] = {d_oo, d_on, d_no, d_nn};       // (rolled loops below: this is cold code, keep it small)
            CorrEv *dst = bd.corr + (size_t)(jt * (jt - 1) / 2 + ju) * CORR_PER_PAIR;
#pragma unroll 1
            for (int gi = 0; gi < FRMC_MAX_GRIDS; ++gi)
#pragma unroll 1
                for (int c4 = 0; c4 < 4; ++c4) {
                    CorrEv ev;
                    ev.sym = -1; ev.ord = 0;
                    if (gi < gs.n) {
                        ev = corr_event(d4[c4], same, slab_, sym, gs.grid[gi], nEl);
                        const bool plus = (c4 == 0 || c4 == 3), redone = (c4 == 1 || c4 == 3);
                        if (ev.ord & 2) ev.ord = (ev.ord & ~6) | (redone ? 2 : 4);
                        if (plus) ev.ord |= 1;
                    }
                    dst[gi * 4 + c4] = ev;
                }
        }
    }
    grid_arrive(bars);
    bar_target += gridDim.x;
    grid_wait(bars, bar_target);
    BATCH_STAMP(1);
    // ---- (2) delta pass of the whole batch
    {
        const int T = gridDim.x * blockDim.x;
        const float4 padrec = make_float4(0.f, 0.f, 0.f, __uint_as_float(PAD_META));
        // A small store leaves most of the grid idle while each busy thread walks all the moved atoms one after the other
        // (10^4 records against 64 moved atoms: 2 warps per CTA, 64 distance pairs each).  When the records fill less than
        // half of the grid, S copies of the pass run side by side, copy s taking the moved atoms t with t % S == s.
        const int span = (npad + 31) & ~31;
        const int fit = T / span;
        const int S = (fit >= 64) ? 64 : (fit >= 32) ? 32 : (fit >= 16) ? 16 : (fit >= 8) ? 8 : (fit >= 4) ? 4 : (fit >= 2) ? 2 : 1;
        const int gid = blockIdx.x * blockDim.x + tid;
        const int sub = (S > 1) ? gid / span : 0;
        const int pbase = (S > 1) ? ((sub < S) ? gid - sub * span : npad) : gid;
        const unsigned long long every = (S == 1) ? ~0ull : (S == 2) ? 0x5555555555555555ull : (S == 4) ? 0x1111111111111111ull :
                                         (S == 8) ? 0x0101010101010101ull : (S == 16) ? 0x0001000100010001ull :
                                         (S == 32) ? 0x0000000100000001ull : 1ull;
        const unsigned long long sub_mask = every << (sub & (S - 1));
        float4 nxt[DELTA_UNROLL];
        {
            const int p0 = pbase;
#pragma unroll
            for (int u = 0; u < DELTA_UNROLL; ++u) { const int p = p0 + u * T; nxt[u] = (p < npad) ? __ldcg(atoms + p) : padrec; }
        }
        for (int p0 = pbase; p0 < npad; p0 += DELTA_UNROLL * T) {
            float4 a[DELTA_UNROLL];
#pragma unroll
            for (int u = 0; u < DELTA_UNROLL; ++u) a[u] = nxt[u];
            {
                const int q0 = p0 + DELTA_UNROLL * T;
#pragma unroll
                for (int u = 0; u < DELTA_UNROLL; ++u) { const int p = q0 + u * T; nxt[u] = (p < npad) ? __ldcg(atoms + p) : padrec; }
            }
#pragma unroll
            for (int u = 0; u < DELTA_UNROLL; ++u) {
                const int p = p0 + u * T;
                const uint32_t mj = __float_as_uint(a[u].w);
                // the warp holds one aligned 32-record sub-block: its bounding box from the records themselves
                // (so it is never stale), then lane l tests moved atoms l, l+32 against it; only the moved atoms
                // whose old or new position can reach the box are swept (exact: blocks_far bounds the computed d^2
                // from below, common.cuh)
                unsigned long long near_mask;
                if (cp.enabled) {
                    // min / max over the warp as one REDUX each, on the order-preserving integer image of the floats
                    const float v[3] = {a[u].x, a[u].y, a[u].z};
                    const bool fin = mj != PAD_META && isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]);
                    float lo[3], hi[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float f = cp.pbc ? (v[c] - floorf(v[c])) : v[c];
                        const unsigned int kmin = __reduce_min_sync(0xffffffffu, fin ? float_order(f) : 0xFFFFFFFFu);
                        const unsigned int kmax = __reduce_max_sync(0xffffffffu, fin ? float_order(f) : 0u);
                        lo[c] = order_float(kmin); hi[c] = order_float(kmax);
                    }
                    // (w of the low corner: the rounding margin of fl(xi - xj), here from the store-wide coordinate bound)
                    const float4 loJ = make_float4(lo[0], lo[1], lo[2], bd.box_eps);
                    const float4 hiJ = make_float4(hi[0], hi[1], hi[2], (lo[0] <= hi[0]) ? 0.f : 1.f);
                    near_mask = 0ull;
                    for (int t0 = 0; t0 < na; t0 += 32) {
                        const int t = t0 + (tid & 31);
                        // the moved atom's own box spans its old and its new position (one test for both)
                        const bool reach = (t < na) && !blocks_far(bs.fOld[t], bs.fNew[t], loJ, hiJ, cp);
                        near_mask |= (unsigned long long)__ballot_sync(0xffffffffu, reach) << t0;
                    }
                } else {
                    near_mask = (na >= 64) ? ~0ull : ((1ull << na) - 1ull);
                }
                if (mj == PAD_META || near_mask == 0ull) continue;
                bool moved_here = false;                 // this record is one of the batch's moved atoms (rare; its
                                                         // own old position lies in the box, so it is in near_mask)
                for (unsigned long long mk = near_mask; mk; mk &= mk - 1ull) moved_here |= (bs.sPos[__ffsll((long long)mk) - 1] == p);
                near_mask &= sub_mask;                   // this copy's share of the moved atoms
                const int ej = mj & 0xFF;
                for (unsigned long long mk = near_mask; mk; mk &= mk - 1ull) {
                    const int t = __ffsll((long long)mk) - 1;
                    const int j = bs.sProp[t];
                    if (moved_here) {
                        // a proposal never pairs its atoms with the stored copy of its own atoms (handled below)
                        bool own = false;
                        for (int v = bs.in_first[j]; v < bs.in_first[j + 1]; ++v) own |= (bs.sPos[v] == p);
                        if (own) continue;
                    }
                    const float4 o = bs.sOld[t], nw = bs.sNew[t];
                    const uint32_t mt = __float_as_uint(o.w);
                    const float d2o = dist2<MODE>(o.x, o.y, o.z, a[u].x, a[u].y, a[u].z, L);
                    const float d2n = dist2<MODE>(nw.x, nw.y, nw.z, a[u].x, a[u].y, a[u].z, L);
                    const bool ho = (d2o >= gs.t2lo) && (d2o < gs.t2hi);
                    const bool hn = (d2n >= gs.t2lo) && (d2n < gs.t2hi);
                    if (ho || hn) {
                        const int same = (mt >> 8) == (mj >> 8);
                        const int et = (int)(mt & 0xFF);
                        const int slab_ = et * nEl + ej;
                        const int sym = sym_index(et, ej, nEl);
                        unsigned long long ov = 0;
                        if (ho) batch_hit(d2o, -1, same, slab_, sym, j, gs, bd, nEl, ov);
                        if (hn) batch_hit(d2n, +1, same, slab_, sym, j, gs, bd, nEl, ov);
                        if (ov) atomicAdd(&bd.bov[j], ov);
                    }
                }
            }
        }
        if (blockIdx.x == 0) {
            // pairs inside a proposal's group: (t,u), u earlier in the list -> slab [el_t, el_u] (the M-F convention)
            for (int e = tid; e < na * na; e += blockDim.x) {
                const int t = e / na, u = e - t * na;
                if (u >= t || bs.sProp[t] != bs.sProp[u]) continue;
                const float4 ot = bs.sOld[t], ou = bs.sOld[u], nt = bs.sNew[t], nu = bs.sNew[u];
                const uint32_t mt = __float_as_uint(ot.w), mu = __float_as_uint(ou.w);
                const int same = (mt >> 8) == (mu >> 8);
                const int et = (int)(mt & 0xFF), eu = (int)(mu & 0xFF);
                const int slab_ = et * nEl + eu;
                const int sym = sym_index(et, eu, nEl);
                const float d2o = dist2<MODE>(ot.x, ot.y, ot.z, ou.x, ou.y, ou.z, L);
                const float d2n = dist2<MODE>(nt.x, nt.y, nt.z, nu.x, nu.y, nu.z, L);
                unsigned long long ov = 0;
                if ((d2o >= gs.t2lo) && (d2o < gs.t2hi)) batch_hit(d2o, -1, same, slab_, sym, bs.sProp[t], gs, bd, nEl, ov);
                if ((d2n >= gs.t2lo) && (d2n < gs.t2hi)) batch_hit(d2n, +1, same, slab_, sym, bs.sProp[t], gs, bd, nEl, ov);
                if (ov) atomicAdd(&bd.bov[bs.sProp[t]], ov);
            }
        }
    }
    grid_arrive(bars);
    bar_target += gridDim.x;
    grid_wait(bars, bar_target);
    BATCH_STAMP(2);
    // ---- (3) rounds
    int cur = 0, ri = __ldcg(&bd.run->n_rand), n_acc = 0, rounds = 0;
    const int acc_before = __ldcg(&bd.run->n_accepted);
    float total = __ldcg(&bd.run->total);
    unsigned int acc_mask = 0u;
    bool stopped = false;
    const int lane = tid & 31, wrp = tid >> 5;
    bool warm = false;                                   // this CTA's epilogue tables are staged
    // Lazy commit of the symmetrised totals.  They exist twice (GridDev::tot / ::stot, equal between launches): buffer
    // `vis` is complete and visible to every CTA; the acceptances of a round are folded into the OTHER buffer without
    // a barrier, while the next round already evaluates on buffer vis + the deltas of those acceptances (pend).  The
    // barrier that ends that round's evaluations makes the other buffer the visible one.  A barrier right after the
    // commit is only needed when an accepted proposal pairs with an atom of an unresolved one (its delta is corrected).
    int vis = 0;
    unsigned int pend = 0u;
    bool inflight = false, other_stale = false;
    // ---- the plan of a round (thread 0).  Which G nodes are worth evaluating is a PREDICTION; whatever it is, a
    // decision is only ever taken on a node whose assumed set equals the set the walk has accepted, so the outcome is
    // exact.  Two proposals change the total standard error almost additively (the cross term is second order in
    // 1/N), so the change a proposal made on one state (est) predicts its fate on the states that follow:
    //   1. a chain along the predictions from proposal c on, the assumed set growing with every predicted acceptance
    //      (at most BATCH_MAX_SPEC); a node needs that no assumed proposal shares an atom (in.share) or an in-range
    //      pair (bs.near) with its proposal -- then the proposal's delta needs no correction;
    //   2. the first proposal without a prediction on the chain's state, and the proposals behind it on the committed
    //      state: they are the next rounds' predictions, and while nothing is assumed yet they extend the chain
    //      (the "all rejected so far" chain of a run without predictions).
    // Built by warp 0, lane j = proposal j (a launch has at most 32): everything is a vote or a prefix count.
    // Refit schedules: an evaluation made at accepted count n refits when n % frequency == 0, and an accepted refit
    // changes the model's scale factor for everything behind it.  A node may therefore only assume acceptances whose
    // own evaluations do not refit: counts base .. base + |A| - 1 must not hit a multiple of any frequency.
    bool any_freq = false;
    for (int mm = 0; mm < FRMC_MAX_MODELS; ++mm) any_freq = any_freq || bd.freq[mm] > 0;
    auto spec_limit = [&](unsigned int accm) {
        int lim = BATCH_MAX_SPEC;
        if (fly && gs.n >= 3) lim = (gs.n == 3) ? 4 : 3;     // the commit's event list: pairs x 4 x grids <= CORR_COMMIT_MAX
        if (!any_freq) return lim;
        const unsigned int acc = (unsigned int)__popc(accm);
        for (int mm = 0; mm < ms.n; ++mm)
            if (bd.freq[mm] > 0) {
                const unsigned int f = (unsigned int)bd.freq[mm];
                const unsigned int t = (f - (bs.s_cnt_mod[mm] + acc) % f) % f;
                lim = min(lim, (int)t);
            }
        return lim;
    };
    auto build_plan = [&](int c, unsigned int accm, int ri_off) {
        const unsigned int FULL = 0xFFFFFFFFu;
        const int max_spec = spec_limit(accm);
        const int j = lane;
        const unsigned int below = (1u << j) - 1u, from_c = ~((1u << c) - 1u);
        const unsigned int he = bs.has_est;
        const bool in_rng = j >= c && j < np;
        const bool has = in_rng && ((he >> j) & 1u);
        const unsigned int has_m = __ballot_sync(FULL, has), rng_m = __ballot_sync(FULL, in_rng);
        const unsigned int noest = rng_m & ~has_m;
        const int f = noest ? __ffs(noest) - 1 : np;                 // first proposal without a prediction
        const bool lead = in_rng && j < f;                           // the leading run of predicted proposals
        const bool worse = lead && bs.est[j] > 0.0f;
        const unsigned int worse_m = __ballot_sync(FULL, worse);
        int dec = 1;
        if (worse) {                                                 // (lanes beyond the round's G nodes are never used)
            const float u = bd.rand_per_proposal ? bs.s_prand[j] : bs.s_rand[min(ri_off + __popc(worse_m & below), 2 * BATCH_MAX_GROUPS - 1)];
            dec = (u > bd.tol) ? 0 : 1;
        }
        const unsigned int acc_pred = __ballot_sync(FULL, lead && dec);
        const unsigned int A_j = acc_pred & below & from_c;         // predicted acceptances in front of j
        const unsigned int sh_j = in_rng ? bs.in_share[j] : 0u, nr_j = in_rng ? bs.near[j] : 0u;
        const bool fits = !((sh_j & (accm | A_j)) || (!fly && (nr_j & A_j))) && __popc(A_j) <= max_spec;
        const unsigned int bad = __ballot_sync(FULL, lead && !fits);
        int chain_end = bad ? __ffs(bad) - 1 : f;
        if (chain_end - c > G) chain_end = c + G;
        const int L = chain_end - c;
        // 1. the chain along the predictions
        if (in_rng && j < chain_end) {
            const int sl = j - c;
            bs.slot_k[sl] = j; bs.slot_A[sl] = A_j;
            bs.slot_child[sl][dec ^ 1] = -1;
            bs.slot_child[sl][dec] = (j + 1 < chain_end) ? sl + 1 : -1;
        }
        __syncwarp();
        int n = L;
        // 2. the first proposal without a prediction, on the chain's state
        const unsigned int A_f = acc_pred & from_c & ((f < 32) ? ((1u << f) - 1u) : FULL);
        const unsigned int sh_f = __shfl_sync(FULL, sh_j, f & 31), nr_f = __shfl_sync(FULL, nr_j, f & 31);
        const bool open = (chain_end == f) && (f < np) && (L < G) &&
                          !((sh_f & (accm | A_f)) || (!fly && (nr_f & A_f))) && __popc(A_f) <= max_spec;
        const int dec_last = __shfl_sync(FULL, dec, (f - 1) & 31);   // predicted decision of the chain's last proposal
        int est_from = chain_end;
        if (open) {
            if (j == 0) {
                bs.slot_k[n] = f; bs.slot_A[n] = A_f; bs.slot_child[n][0] = -1; bs.slot_child[n][1] = -1;
                if (L > 0) bs.slot_child[n - 1][dec_last] = n;
            }
            ++n;
            est_from = f + 1;
        }
        // 3. behind it, on the committed state: the next rounds' predictions.  While nothing is assumed they also
        //    extend the chain ("all rejected so far").
        const unsigned int conflict = __ballot_sync(FULL, in_rng && j >= est_from && (sh_j & accm));
        const int stop_at = conflict ? __ffs(conflict) - 1 : np;
        const bool cand = in_rng && j >= est_from && j < stop_at && !has;
        const unsigned int cand_m = __ballot_sync(FULL, cand);
        const int sl = n + __popc(cand_m & below);
        __syncwarp();
        if (cand && sl < G) {
            bs.slot_k[sl] = j; bs.slot_A[sl] = 0u; bs.slot_child[sl][0] = -1; bs.slot_child[sl][1] = -1;
        }
        __syncwarp();
        if (open && A_f == 0u) {
            // proposals f+1, f+2, ... as long as they are consecutive candidates: rejection children of one another
            const unsigned int after_f = (f + 1 < 32) ? ~((1u << (f + 1)) - 1u) : 0u;
            const unsigned int gap = ~cand_m & after_f;
            const int run_end = gap ? __ffs(gap) - 1 : 32;             // exclusive
            if (cand && j < run_end && sl < G) bs.slot_child[sl - 1][0] = sl;
        }
        if (j == 0) {
            bs.n_slots = min(G, n + __popc(cand_m));
            int P = L;
            if (open) {
                P = L + 1;
                if (A_f == 0u) {
                    const unsigned int after_f = (f + 1 < 32) ? ~((1u << (f + 1)) - 1u) : 0u;
                    const unsigned int gap = ~cand_m & after_f;
                    const unsigned int run = cand_m & after_f & (gap ? ((1u << (__ffs(gap) - 1)) - 1u) : 0xFFFFFFFFu);
                    P += min(__popc(run), G - P);
                }
            }
            bs.path_len = P;
            bs.path_exp = (L > 0) ? ((acc_pred >> c) & ((L < 32) ? ((1u << L) - 1u) : 0xFFFFFFFFu)) : 0u;
        }
        __syncwarp();
    };
    if (tid == 0) bs.has_est = 0u;
    __syncthreads();
    if (wrp == 0) build_plan(0, 0u, 0);
    __syncthreads();
    while (cur < np && !stopped) {
        ++rounds;
        unsigned long long t_round = 0;
        if (stamps && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_round));
        const int ns = bs.n_slots;
        const int par = rounds & 1;                          // slot buffers alternate: a CTA may still read round r-1's
        if (epi && group < ns) {
            const int k = bs.slot_k[group];
            unsigned int A = bs.slot_A[group];
            const int sb = par * BATCH_MAX_GROUPS + group;
            EpiOut eo;
            const GridDev &Gd = gs.grid[ms.m[m].grid];
            const long long per = (long long)Gd.nsym * Gd.g.hs;
            eo.base = vis ? Gd.stot : Gd.tot;
            eo.adds[0] = bd.bsym[ms.m[m].grid] + (long long)k * per;
            eo.amask[0] = bs.symmask[k];
            eo.n_adds = 1;
            for (A |= pend; A; A &= A - 1u) {                // assumed acceptances and those not yet in the visible buffer
                const int a = __ffs(A) - 1;
                eo.adds[eo.n_adds] = bd.bsym[ms.m[m].grid] + (long long)a * per; eo.amask[eo.n_adds] = bs.symmask[a];
                ++eo.n_adds;
            }
            eo.total = bd.btotal[m] + (long long)sb * ms.m[m].n_out;
            eo.res = bd.res + sb * 2 * FRMC_MAX_MODELS;
            eo.terms = ((bd.defer_mask >> m) & 1u) ? bd.bterm[m] + (long long)sb * ms.m[m].n_out : nullptr;
            eo.warm = warm ? 1 : 0;
            // on-the-fly pair corrections: every pair (x, y) with x in A + {k}, y in A, y < x, atoms in range
            eo.ev = bs.ev; eo.ev_bits = bs.ev_bits; eo.n_ev = 0;
            {
                const unsigned int Aas = bs.slot_A[group];
                bool any_pair = false;                        // (uniform) does any member of A + {k} see an earlier member of A?
                if (fly && Aas) {
                    for (unsigned int xm = Aas | (1u << k); xm; xm &= xm - 1u) {
                        const int x = __ffs(xm) - 1;
                        any_pair = any_pair || (bs.near[x] & Aas & ((1u << x) - 1u)) != 0u;
                    }
                }
                if (any_pair) {
                    __syncthreads();                          // the previous use of the list (this CTA's last node / commit) is over
                    if (tid < 32) bs.ev_bits[tid] = 0u;
                    if (tid == 0) bs.n_ev = 0;
                    __syncthreads();
                    const int gi = ms.m[m].grid;
                    // thread = (x slot, y, combo): x runs over the members of A and k itself
                    const unsigned int X = Aas | (1u << k);
                    const int nx = __popc(X);
                    for (int e = tid; e < nx * 32 * 4; e += blockDim.x) {
                        const int xi = e >> 7, y = (e >> 2) & 31, c4 = e & 3;
                        int x = 0;
                        { unsigned int xm = X; for (int q = 0; q < xi; ++q) xm &= xm - 1u; x = __ffs(xm) - 1; }
                        if (y >= x || !((Aas >> y) & 1u) || !((bs.near[x] >> y) & 1u)) continue;
                        const int2 raw = __ldcg(reinterpret_cast<const int2 *>(bd.corr + ((size_t)(x * (x - 1) / 2 + y) * CORR_PER_PAIR + gi * 4 + c4)));
                        CorrEv cv;
                        cv.sym = raw.x; cv.ord = raw.y;
                        if (cv.sym < 0) continue;
                        const int at = atomicAdd(&bs.n_ev, 1);
                        if (at < CORR_NODE_MAX) {
                            bs.ev[at] = (unsigned int)cv.sym | ((cv.ord & 1) ? 0u : 0x80000000u);
                            atomicOr(&bs.ev_bits[((cv.sym & 0xFFFF) & 1023) >> 5], 1u << (cv.sym & 31));
                        }
                    }
                    __syncthreads();
                    if (bs.n_ev > CORR_NODE_MAX) __trap();   // cannot happen: <= 15 pairs x 4 events
                    eo.n_ev = bs.n_ev;
                }
            }
            eo.refit = 0;
            if (bd.freq[m] > 0) {
                const unsigned int f = (unsigned int)bd.freq[m];
                eo.refit = ((bs.s_cnt_mod[m] + (unsigned int)__popc(acc_mask) + (unsigned int)__popc(bs.slot_A[group])) % f == 0u) ? 1 : 0;
            }
            eo.scale = bs.s_csf[m];
            long long *est = (stamps && blockIdx.x == 1 && rounds <= 64) ? stamps + BATCH_STAMP_SLOTS + (rounds - 1) * 128 : nullptr;
            if (est && tid == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); est[120] = (long long)gt_; est[121] = (long long)t_round; }
            epilogue_run<true>(es, epi_smem, ms, gs, m, slab, nullptr, nullptr, nullptr, bd.tickets + sb * FRMC_MAX_MODELS, est, eo);
            warm = true;
        }
        BATCH_STAMP(4 + 5 * (rounds - 1) + 0);           // CTA 0's own epilogue done
        grid_arrive(bars);
        bar_target += gridDim.x;
        grid_wait(bars, bar_target);
        BATCH_STAMP(4 + 5 * (rounds - 1) + 1);           // every epilogue of the round done
        if (inflight) { vis ^= 1; pend = 0u; inflight = false; other_stale = true; }   // the last commit is visible now
        // decisions: every CTA walks the same numbers down the tree to the same conclusion
        {
            // chi2 of every slot: models in defer_mask left their terms (all slab CTAs wrote a slice) and one warp per
            // (slot, model) sums them in numpy's pairwise order, straight from L2; the other models left their chi2
            const int n_def = __popc(bd.defer_mask);
            const int n_tasks = ns * n_def;
            if (tid >= EPI_THREADS - 32) bs.s_rand[lane] = __ldcg(bd.rand + ri + lane);
            if (tid >= EPI_THREADS - 64 && tid < EPI_THREADS - 32) {
                for (int e = lane; e < ns * ms.n; e += 32) {
                    const int sl = e / ms.n, mm = e - sl * ms.n;
                    if (!((bd.defer_mask >> mm) & 1u)) bs.s_chi[sl][mm] = __ldcg(bd.res + (par * BATCH_MAX_GROUPS + sl) * 2 * FRMC_MAX_MODELS + mm);
                }
            }
            for (int task = wrp; task < n_tasks; task += EPI_THREADS / 32) {
                const int sl = task / n_def;
                int mm = 0;
                for (int c = task - sl * n_def, x = 0; x < ms.n; ++x)
                    if ((bd.defer_mask >> x) & 1u) { if (c == 0) { mm = x; break; } --c; }
                const float chi = warp_pairwise_sum(bd.bterm[mm] + (long long)(par * BATCH_MAX_GROUPS + sl) * ms.m[mm].n_out, bs.sched[mm],
                                                    ms.m[mm].pw_leaves, bs.pw_scratch[wrp]);
                if (lane == 0) bs.s_chi[sl][mm] = chi;
            }
            __syncthreads();
            // the engine's total standard error of a slot: np.sum of the float32 list [chi2_m / varianceSquared_m]
            // (Engine.py:3024-3029; sequential for fewer than 8 terms, numpy's 8-accumulator tree for exactly 8)
            if (tid < ns) {
                float term[FRMC_MAX_MODELS];
#pragma unroll
                for (int i = 0; i < FRMC_MAX_MODELS; ++i) term[i] = (i < ms.n) ? __fdiv_rn(bs.s_chi[tid][i], bd.var2[i]) : 0.0f;
                float tot;
                if (ms.n == 8) {
                    tot = __fadd_rn(__fadd_rn(__fadd_rn(term[0], term[1]), __fadd_rn(term[2], term[3])),
                                    __fadd_rn(__fadd_rn(term[4], term[5]), __fadd_rn(term[6], term[7])));
                } else {
                    tot = 0.0f;
#pragma unroll
                    for (int i = 0; i < FRMC_MAX_MODELS; ++i) if (i < ms.n) tot = __fadd_rn(tot, term[i]);
                }
                bs.s_pt[tid] = tot;
                // a node on the committed state predicts its proposal's fate in the rounds to come
                if (bs.slot_A[tid] == 0u) { bs.est[bs.slot_k[tid]] = __fsub_rn(tot, total); atomicOr(&bs.has_est, 1u << bs.slot_k[tid]); }
            }
        }
        __syncthreads();
        if (wrp == 0) {
            {
                // lane s = slot s of the linked path.  Each lane decides its proposal as if every slot in front of it
                // had gone the expected way; the walk ends at the first slot that did not (its own decision stands: its
                // node is still the true state).  A slot that is not exactly the expected state cuts the path.
                const unsigned int FULL = 0xFFFFFFFFu, below = (1u << lane) - 1u;
                const unsigned int pexp = bs.path_exp;
                int P = bs.path_len;
                const bool on = lane < P;
                const bool exact = on && bs.slot_k[lane] == cur + lane && bs.slot_A[lane] == ((pexp & below) << cur);
                const unsigned int inexact = __ballot_sync(FULL, on && !exact);
                if (inexact) P = __ffs(inexact) - 1;                       // (slot 0 is (cur, nothing) by construction)
                if (P < 1) __trap();                                       // never spin without progress
                const float nt = (lane < P) ? bs.s_pt[lane] : 0.0f;
                const unsigned int acc_before = pexp & below;
                const float nt_prev = __shfl_sync(FULL, nt, acc_before ? 31 - __clz(acc_before) : 0);
                const float tl_s = acc_before ? nt_prev : total;
                const bool worse = lane < P && nt > tl_s;
                const unsigned int worse_m = __ballot_sync(FULL, worse);
                int dec = 1;
                if (worse) {
                    const float u = bd.rand_per_proposal ? bs.s_prand[min(cur + lane, BATCH_MAX_PROPS - 1)] : bs.s_rand[__popc(worse_m & below)];
                    dec = (u > bd.tol) ? 0 : 2;
                }
                const bool surprise = lane < P - 1 && ((dec != 0) != (((pexp >> lane) & 1u) != 0u));
                const unsigned int sur = __ballot_sync(FULL, surprise);
                const int s_end = sur ? __ffs(sur) - 1 : P - 1;            // last slot the walk resolves
                const bool res = lane <= s_end;
                const unsigned int acc_slots = __ballot_sync(FULL, res && dec != 0);
                if (res) { bs.path_slot[lane] = lane; bs.path_dec[lane] = dec; }
                const int last = acc_slots ? 31 - __clz(acc_slots) : -1;
                const float tl = __shfl_sync(FULL, nt, last < 0 ? 0 : last);
                if (lane == 0) {
                    const int k = cur + s_end + 1;
                    const unsigned int A = acc_slots << cur;
                    const int used = __popc(worse_m & ((s_end < 31) ? ((2u << s_end) - 1u) : FULL));
                    // a proposal that moves an atom an accepted proposal of this launch has moved ends the launch
                    const bool stop = (k < np) && (bs.in_share[k] & (acc_mask | A));
                    bs.s_acc = A; bs.s_last = last; bs.s_cur = k; bs.s_ri = ri + used; bs.s_stopped = stop ? 1 : 0;
                    bs.s_total = last < 0 ? total : tl;
                }
            }
            __syncwarp();
            const int k = bs.s_cur;
            const unsigned int A = bs.s_acc;
            if (k < np && !bs.s_stopped) {
                // the next round's plan.  (The deltas of unresolved proposals next to an accepted one are about to be
                // corrected by a few pair events; their predictions stay, they only steer the plan.)
                build_plan(k, acc_mask | A, bs.s_ri - ri);
            }
        }
        __syncthreads();
        if (blockIdx.x == 0) {
            // decisions and chi2 of the proposals resolved in this round
            const int n_path = bs.s_cur - cur, per = ms.n + 1;
            for (int e = tid; e < n_path * per; e += blockDim.x) {
                const int i = e / per, mm = e - i * per;
                if (mm == ms.n) bd.out_dec[in.out_base + cur + i] = bs.path_dec[i];
                else bd.out_chi2[(long long)(in.out_base + cur + i) * ms.n + mm] = bs.s_chi[bs.path_slot[i]][mm];
            }
        }
        const unsigned int Aset = bs.s_acc;
        const int last = bs.s_last;
        cur = bs.s_cur; ri = bs.s_ri; stopped = bs.s_stopped != 0; total = bs.s_total;
        if (Aset && tid < ms.n && bd.freq[tid] > 0)     // accept_move keeps the scale factor the accepted evaluation used
            bs.s_csf[tid] = __ldcg(bd.res + (par * BATCH_MAX_GROUPS + last) * 2 * FRMC_MAX_MODELS + FRMC_MAX_MODELS + tid);
        __syncthreads();
        BATCH_STAMP(4 + 5 * (rounds - 1) + 2);           // decisions made
        if (Aset) {
            // ---- commit the accepted proposals (all CTAs), correct the deltas of the unresolved proposals
            acc_mask |= Aset;
            n_acc += __popc(Aset);
            int acc_j[BATCH_MAX_SPEC + 1], n_aj = 0;
            for (unsigned int a = Aset; a; a &= a - 1u) acc_j[n_aj++] = __ffs(a) - 1;
            const bool last_round = (cur >= np) || stopped;
            // does an accepted proposal pair with an atom of an unresolved one?  (every warp finds the same answer)
            bool need_bar = __ballot_sync(0xFFFFFFFFu, lane >= cur && lane < np && (bs.near[lane] & Aset)) != 0u;
            // on-the-fly mode: several acceptances of this round may pair with EACH OTHER (their nodes saw the events of
            // those pairs from the table); the commit adds the same events to the totals and the ordered counts.  Every
            // CTA gathers the list; entries: sym = grid << 24 | row << 16 | bin, ord = grid << 28 | cell << 3 | codes.
            CorrEv *clist = reinterpret_cast<CorrEv *>(&bs.pw_scratch[0][0]);      // idle outside the decision phase
            int n_cev = 0;
            bool acc_pairs = false;                           // (uniform) do two acceptances of this round see each other?
            if (fly && (Aset & (Aset - 1u)))
                for (int x = 0; x < n_aj; ++x) acc_pairs = acc_pairs || (bs.near[acc_j[x]] & Aset & ((1u << acc_j[x]) - 1u)) != 0u;
            if (acc_pairs && !(bd.debug & 2)) {
                if (tid < 32) bs.ev_bits[tid] = 0u;
                if (tid == 0) bs.n_cev = 0;
                __syncthreads();
                const int per_x = 32 * gs.n * 4;
                for (int e = tid; e < n_aj * per_x; e += blockDim.x) {
                    const int x = acc_j[e / per_x], rem = e % per_x, y = rem / (gs.n * 4), g4 = rem % (gs.n * 4), gi = g4 >> 2;
                    if (y >= x || !((Aset >> y) & 1u) || !((bs.near[x] >> y) & 1u)) continue;
                    const int2 raw = __ldcg(reinterpret_cast<const int2 *>(bd.corr + ((size_t)(x * (x - 1) / 2 + y) * CORR_PER_PAIR + g4)));
                    if (blockIdx.x == 0 && (raw.y & 6))          // edge-overflow events follow the pair they belong to
                        atomicAdd(&bd.bov[x], (raw.y & 2) ? 1ull : ~0ull);
                    if (raw.x < 0) continue;
                    const int at = atomicAdd(&bs.n_cev, 1);
                    if (at < CORR_COMMIT_MAX) {
                        clist[at].sym = raw.x | (gi << 24);
                        clist[at].ord = raw.y | (gi << 28);
                        const int hs_g = gs.grid[gi].g.hs;
                        const int csym = ((raw.x >> 16) & 0xFF) * hs_g + (raw.x & 0xFFFF);
                        atomicOr(&bs.ev_bits[((csym + 977 * gi) & 1023) >> 5], 1u << ((csym + 977 * gi) & 31));
                        const int cord = (raw.y >> 3) & 0x1FFFFFF;
                        atomicOr(&bs.ev_bits[((cord + 331 + 977 * gi) & 1023) >> 5], 1u << ((cord + 331 + 977 * gi) & 31));
                    }
                }
                __syncthreads();
                n_cev = bs.n_cev;
                if (n_cev > CORR_COMMIT_MAX) __trap();       // spec_limit keeps pairs x 4 x grids within the list
                if (n_cev > 0) need_bar = true;               // the lazy commit's "pending" set assumes non-interacting acceptances
            }
            const long long stride = (long long)gridDim.x * blockDim.x;
            const long long gt = (long long)blockIdx.x * blockDim.x + tid;
            const int lb = par * BATCH_MAX_GROUPS + last;    // the evaluation that saw all of them
            // one list of items over all grids and models (symmetrised totals | ordered counts | model totals), so a
            // thread usually has one item and the whole commit is a single L2 round trip
            long long n_items_c = 0;
            for (int gi = 0; gi < gs.n; ++gi) n_items_c += (long long)gs.grid[gi].nsym * gs.grid[gi].g.hs + 2 * gs.grid[gi].cells;
            for (int mm = 0; mm < ms.n; ++mm) n_items_c += ms.m[mm].n_out;
            for (long long it0 = gt; it0 < ((bd.debug & 4) ? 0 : n_items_c); it0 += stride) {
                long long c = it0;
                bool done = false;
                for (int gi = 0; gi < gs.n && !done; ++gi) {
                    const GridDev &Gd = gs.grid[gi];
                    const long long ns_ = (long long)Gd.nsym * Gd.g.hs;
                    if (c < ns_) {
                        int *bv = vis ? Gd.stot : Gd.tot, *bo = vis ? Gd.tot : Gd.stot;
                        const int t0 = __ldcg(bv + c);
                        int v = 0;
#pragma unroll
                        for (int x = 0; x <= BATCH_MAX_SPEC; ++x) if (x < n_aj) v += __ldcg(bd.bsym[gi] + (long long)acc_j[x] * ns_ + c);
                        if (n_cev && ((bs.ev_bits[(((int)c + 977 * gi) & 1023) >> 5] >> (((int)c + 977 * gi) & 31)) & 1u)) {
                            const int key = (gi << 24) | ((int)(c / Gd.g.hs) << 16) | (int)(c % Gd.g.hs);
                            for (int e = 0; e < n_cev; ++e) if (clist[e].sym == key) v += (clist[e].ord & 1) ? 1 : -1;
                        }
                        bo[c] = t0 + v;                                  // every cell: the other buffer may be one commit behind
                        if (last_round && v) bv[c] = t0 + v;             // the launch ends with both buffers equal
                        done = true;
                    } else if ((c -= ns_) < 2 * Gd.cells) {
                        const unsigned long long c0 = __ldcg(Gd.counts + c);
                        int d = 0;
#pragma unroll
                        for (int x = 0; x <= BATCH_MAX_SPEC; ++x) if (x < n_aj) d += __ldcg(bd.bdelta[gi] + (long long)acc_j[x] * 2 * Gd.cells + c);
                        if (n_cev && ((bs.ev_bits[(((int)c + 331 + 977 * gi) & 1023) >> 5] >> (((int)c + 331 + 977 * gi) & 31)) & 1u)) {
                            for (int e = 0; e < n_cev; ++e)
                                if (((clist[e].ord >> 3) & 0x1FFFFFF) == (int)c && ((clist[e].ord >> 28) & 3) == gi) d += (clist[e].ord & 1) ? 1 : -1;
                        }
                        if (d) Gd.counts[c] = (unsigned long long)((long long)c0 + d);
                        done = true;
                    } else c -= 2 * Gd.cells;
                }
                for (int mm = 0; mm < ms.n && !done; ++mm) {
                    if (c < ms.m[mm].n_out) { bd.total_committed[mm][c] = __ldcg(bd.btotal[mm] + (long long)lb * ms.m[mm].n_out + c); done = true; }
                    else c -= ms.m[mm].n_out;
                }
            }
            if (blockIdx.x == 0) {
                for (int x = 0; x < n_aj; ++x)
                    for (int t = bs.in_first[acc_j[x]] + tid; t < bs.in_first[acc_j[x] + 1]; t += blockDim.x) {
                        atoms[bs.sPos[t]] = bs.sNew[t];
                        if (GEN && real) real[gen->ridx[t]] = make_float4(gen->mreal[3 * t], gen->mreal[3 * t + 1], gen->mreal[3 * t + 2], 0.f);
                    }
                if (tid < ms.n) {
                    bd.run->cchi2[tid] = bs.s_chi[last][tid];
                    bd.run->csf[tid] = bs.s_csf[tid];
                }
            }
            // pair (t of an unresolved proposal, u of an accepted one): the delta pass paired t with u's OLD position.
            // Proposals resolved in this round need none: their nodes had no such pair (bs.near).
            if (need_bar && !(bd.debug & 1)) {
                // One unit of work = (accepted atom u, unresolved atom t, one of the four old/new combinations).  The units
                // are dealt to lane 0 of consecutive WARPS of the whole grid (unit w -> warp w): a few hundred units finish in
                // the time of one (they used to sit in the first lanes of CTA 0, one after the other).
                const int later0 = bs.in_first[cur];                // atoms are listed in proposal order
                const int n_later = na - later0;
                const long long gw = gt >> 5, n_warps = stride >> 5;
                long long base_u = 0;
                for (int x = 0; x < n_aj; ++x) {
                    const int a0 = bs.in_first[acc_j[x]], a1 = bs.in_first[acc_j[x] + 1];
                    const long long n_units = (long long)n_later * (a1 - a0) * 4;
                    if (lane == 0)
                    for (long long w = gw - base_u % n_warps + ((gw < base_u % n_warps) ? n_warps : 0); w < n_units; w += n_warps) {
                        const int c4 = (int)(w & 3);
                        const long long it = w >> 2;
                        const int t = later0 + (int)(it / (a1 - a0)), u = a0 + (int)(it % (a1 - a0));
                        const int j2 = bs.sProp[t];
                        if (!((bs.near[j2] >> acc_j[x]) & 1u)) continue;         // no pair in range (the common case)
                        const float4 ot = bs.sOld[t], nt = bs.sNew[t], ou = bs.sOld[u], nu = bs.sNew[u];
                        const uint32_t mt = __float_as_uint(ot.w), mu = __float_as_uint(ou.w);
                        const int same = (mt >> 8) == (mu >> 8);
                        const int et = (int)(mt & 0xFF), eu = (int)(mu & 0xFF);
                        const int slab_ = et * nEl + eu;
                        const int sym = sym_index(et, eu, nEl);
                        // c4: 0 (old, old) undo -1 | 1 (old, new) redo -1 | 2 (new, old) undo +1 | 3 (new, new) redo +1
                        const float4 pt = (c4 & 2) ? nt : ot, pu = (c4 & 1) ? nu : ou;
                        const float d2 = dist2<MODE>(pt.x, pt.y, pt.z, pu.x, pu.y, pu.z, L);
                        if (!((d2 >= gs.t2lo) && (d2 < gs.t2hi))) continue;
                        unsigned long long ov = 0;                     // edge-overflow events follow the pairs they belong to
                        batch_hit(d2, (c4 == 0 || c4 == 3) ? +1 : -1, same, slab_, sym, j2, gs, bd, nEl, ov);
                        if (ov) atomicAdd(&bd.bov[j2], (c4 & 1) ? ov : (0ull - ov));   // re-done pairs count, undone ones are taken back (modulo 2^64)
                    }
                    base_u += n_units;
                }
            }
            BATCH_STAMP(4 + 5 * (rounds - 1) + 3);       // CTA 0's share of the commit done
            if (last_round) {
                other_stale = false;                     // both buffers were written
            } else if (need_bar) {
                grid_arrive(bars);
                bar_target += gridDim.x;
                grid_wait(bars, bar_target);
                vis ^= 1; other_stale = true;            // the commit and the corrections are visible to everyone
            } else {
                pend = Aset; inflight = true;            // folded into the other buffer, visible after the next barrier
            }
            BATCH_STAMP(4 + 5 * (rounds - 1) + 4);
        }
    }
    if (other_stale) {
        // the launch ended on a round without acceptance while the other buffer was one commit behind
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (int gi = 0; gi < gs.n; ++gi) {
            const GridDev &Gd = gs.grid[gi];
            const int *bv = vis ? Gd.stot : Gd.tot;
            int *bo = vis ? Gd.tot : Gd.stot;
            for (long long c = (long long)blockIdx.x * blockDim.x + tid; c < (long long)Gd.nsym * Gd.g.hs; c += stride) bo[c] = __ldcg(bv + c);
        }
    }
    BATCH_STAMP(3);
    if (blockIdx.x == 0 && tid == 0) {
        unsigned long long ov = 0;
        for (int j = 0; j < cur; ++j) ov += __ldcg(bd.bov + j);
        if (ov) atomicAdd(overflow, ov);
        BatchRun *r = bd.run;
        r->total = total; r->n_rand = ri; r->n_done = in.out_base + cur; r->n_accepted = __ldcg(&r->n_accepted) + n_acc;
        r->rounds = __ldcg(&r->rounds) + rounds;
        __threadfence();
        r->stopped = stopped ? 1 : 0;
    }
  }
#undef BATCH_STAMP
}
}  // namespace frmc

using namespace frmc;

// ------------------------------------------------------------------ host-side store
struct ModelHost {
    ModelDev dev;                    // device pointers (owned)
    float *total_committed = nullptr;
    int adjust_freq = 0;             // scale-factor refit every `freq` accepted moves (0: never)
    float sf_staged = 1.0f;          // scale factor the last evaluation used
    float *amp_w = nullptr, *amp_D = nullptr, *amp_rD = nullptr, *amp_pref = nullptr;   // constants of the one-atom-fewer evaluation
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
    int64_t n = 0, npad = 0;         // n: atoms the store holds NOW (the engine's relative numbering runs over them)
    int64_t n0 = 0;                  // atoms the layout was built for (the "real" numbering: lay.inv, d_orig, d_mol, h_mol, h_el)
    uint64_t layout_gen = 0;         // counts the layouts of this store: per-position tables of other translation units follow it
    std::vector<int32_t> rel2real;   // relative -> real index once atoms have been removed (empty: identity)
    int amp_rel = -1;                // relative index of the atom of the staged amputation
    int nEl = 0, isPBC = 0;
    Lattice L;
    float lo[3], hi[3];
    std::vector<int32_t> h_mol, h_el;
    int32_t *d_mol = nullptr;                   // molecule index by original atom index (full histogram: exact intra/inter test)
    uint32_t mol_span = 0;                      // HostLayout::mol_span
    float4 *d_atoms = nullptr;
    uint32_t *d_orig = nullptr;
    WorkItem *d_items = nullptr;     // rows of the pair list (I tile x J range of an element pair)
    PairLists lists;                 // device-built surviving block pairs, cut into items
    int n_items = 0, R = 1;
    int items_shard = -1, items_nshards = -1;   // which slice of the row list d_items holds
    int n_pairs = 0;                            // element pairs present in that slice
    HostLayout lay;                  // rec freed after upload; segments + inverse permutation kept
    int *d_next = nullptr;
    unsigned long long *d_overflow = nullptr;   // [0] edge-overflow events, [1] block pairs swept by the last compute_data
    float4 *d_bbox = nullptr;                   // 2 records per SEG_PAD block (full-histogram culling)
    std::vector<GridHost> grids;
    std::vector<ModelHost> models;
    bool models_dirty = true;
    ProposalIn prop_in;              // staged proposal, handed to the delta kernel by value
    Proposal *d_prop = nullptr;
    float *h_chi2 = nullptr;         // pinned, device-visible: [FRMC_MAX_MODELS] chi2, [FRMC_MAX_MODELS] u32 sequence numbers, [FRMC_MAX_MODELS] scale factors used
    volatile unsigned int *h_seq = nullptr;
    unsigned int *d_seq = nullptr;   // [FRMC_MAX_MODELS] launch counters + [FRMC_MAX_MODELS] tickets
    long long *d_stamps = nullptr;   // [FRMC_MAX_MODELS][8] clock64 phase stamps of the last epilogue (debug)
    unsigned int seq_expected = 0;
    size_t epi_smem = 0;
    // fused per-move path (one cooperative launch per proposal)
    bool use_fused = true;
    bool fused_ok = false;           // checked in sync_models: cooperative launch + occupancy + epilogue CTA count
    int pending = 0;                 // deferred resolution of the last proposal: 0 none, 1 accept, 2 reject
    unsigned long long *d_bars = nullptr;
    unsigned long long fused_launches = 0, fused_resolves = 0;
    EpiMap epi_map;
    float prop_lo[3], prop_hi[3];
    int state = 0;                   // 0 idle, 1 proposal staged
    float chi2_staged[FRMC_MAX_MODELS];
    // host-side phase timers of frmc_propose (FRMC_STEP_TIMING=1): launch call, wait for chi2, whole call; printed at destroy
    bool step_timing = false;
    double t_launch = 0, t_wait = 0, t_call = 0;
    unsigned long long n_calls = 0;
    // persistent per-move kernel (propose_loop_kernel)
    bool persist_enabled = false;    // frmc_store_set_persistent
    bool persist_ok = false;         // every model fits the resident-slab epilogue, no refit schedule (sync_models)
    bool persist_running = false;
    int persist_mode = -1;           // geometry mode the running kernel was instantiated for
    HostCmd *h_cmd = nullptr;        // mapped pinned
    DevCmd *d_cmd = nullptr;
    unsigned long long *d_pbars = nullptr;   // the persistent kernel's own two barrier counters
    unsigned int cmd_seq = 0;        // last command number written
    int last_prev = 0;               // resolution carried by the last command (to resend it if the kernel had left)
    unsigned long long persist_launches = 0, persist_cmds = 0;
    // runs of proposals resolved on the device (batch_kernel)
    bool batch_ok = false;           // checked in sync_models: resident S(Q) slabs, no refit schedule, co-residency
    bool batch_ready = false;        // buffers below match the current grids and models
    bool batch_fly = true;           // launches use the on-the-fly pair corrections (batch_prepare: dense enough to pay)
    unsigned int batch_defer_mask = 0u;   // sync_models: models whose chi2 is summed after the grid barrier
    BatchDev bdev;
    std::vector<void *> batch_owned;
    unsigned long long *d_bbars = nullptr;
    float *d_brand = nullptr; size_t brand_cap = 0;
    float *d_bout_chi2 = nullptr; int *d_bout_dec = nullptr; size_t bout_cap = 0;
    BatchRun *h_brun = nullptr;      // pinned
    long long *d_bstamps = nullptr;  // debug timeline of the last batch launch (FRMC_BATCH_STAMPS=1)
    cudaEvent_t bev0 = nullptr, bev1 = nullptr;
    unsigned long long batch_launches = 0, batch_rounds = 0, batch_proposals = 0;
    // device-generated runs (generate_batch_kernel)
    float4 *d_real = nullptr;        // engine.realCoordinates by real index (periodic systems); valid while real_valid
    bool real_valid = false;         // every move accepted since frmc_store_set_real_coords went through a generated run
    float rbasis[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // engine.reciprocalBasisVectors
    int *d_inv = nullptr;            // real index -> position in the sorted store
    int *d_goff = nullptr, *d_gidx = nullptr;        // groups
    int n_groups = 0;
    BatchIn *d_bin = nullptr;        // the generated launch
    BatchIn *d_bins = nullptr;       // host proposals: all batches of a run (frmc_run_batch)
    size_t bins_cap = 0;
    std::vector<BatchIn> h_bins;
    GenOut *d_gen = nullptr;
    int *d_gout = nullptr; size_t gout_cap = 0;      // selected group of every proposal of a call
    bool force_general = false;      // a generated coordinate left the fast-wrap window once: general minimum image from then on
    int gen_win_kind = 0;            // 0: no window yet, 1: fast-wrap window, 2: wide window (general minimum image / non-periodic)
    float gen_win_lo[3] = {0, 0, 0}, gen_win_hi[3] = {0, 0, 0};
    unsigned long long accepted = 0;  // the engine's count of accepted moves (refit schedule, Core/Constraint.py:1418-1422)
    float chi2_committed[FRMC_MAX_MODELS];
    // optional per-kernel timing (CUDA events on the store's stream; bench.py's roofline leg)
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    struct Pending { int which; cudaEvent_t a, b; };
    std::vector<Pending> ev_pending;
    double kernel_ms[4] = {0, 0, 0, 0};
    uint64_t kernel_launches[4] = {0, 0, 0, 0};
};

enum { TIME_DELTA = 0, TIME_FULL = 1, TIME_EPILOGUE = 2, TIME_COMMIT = 3 };

// position in the sorted store of the atom the engine calls `idx` (its RELATIVE index: after atoms were removed
// the engine's arrays are np.delete'd, Engine.py:781-788, and every later atom moves down by one)
static inline int pos_of(const frmc_store *s, int idx)
{
    return s->lay.inv[s->rel2real.empty() ? idx : s->rel2real[idx]];
}

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

// persistent per-move kernel lifecycle (defined further down)
static void write_cmd(frmc_store *s, int op, int prev);
static int stop_persistent(frmc_store *s);
static int launch_persistent(frmc_store *s, int mode);

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
    int rc = build_layout(coords, s->n, s->h_mol.data(), s->h_el.data(), s->nEl, s->isPBC, s->lay);
    if (rc) return rc;
    s->npad = s->lay.npad;
    s->mol_span = s->lay.mol_span;
    ++s->layout_gen;
    if (!s->d_mol && s->n > 0) {
        FRMC_CUDA(cudaMalloc(&s->d_mol, sizeof(int32_t) * (size_t)s->n));
        FRMC_CUDA(cudaMemcpyAsync(s->d_mol, s->h_mol.data(), sizeof(int32_t) * (size_t)s->n, cudaMemcpyHostToDevice, s->stream));
    }
    for (int c = 0; c < 3; ++c) { s->lo[c] = s->lay.lo[c]; s->hi[c] = s->lay.hi[c]; }
    if (!s->d_atoms) {
        FRMC_CUDA(cudaMalloc(&s->d_atoms, sizeof(float4) * std::max<int64_t>(s->npad, 1)));
        FRMC_CUDA(cudaMalloc(&s->d_orig, sizeof(uint32_t) * std::max<int64_t>(s->npad, 1)));
        FRMC_CUDA(cudaMalloc(&s->d_bbox, sizeof(float4) * 18 * (size_t)(s->npad / SEG_PAD + 1)));
    }
    if (s->npad > 0) {
        FRMC_CUDA(cudaMemcpyAsync(s->d_atoms, s->lay.rec.data(), sizeof(float4) * s->npad, cudaMemcpyHostToDevice, s->stream));
        FRMC_CUDA(cudaMemcpyAsync(s->d_orig, s->lay.orig.data(), sizeof(uint32_t) * s->npad, cudaMemcpyHostToDevice, s->stream));
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
    build_rows(s->lay, 1, shard, nshards, items);
    std::vector<unsigned char> blob;
    pack_rows(items, blob, s->n_pairs);
    if (s->d_items) { cudaFree(s->d_items); s->d_items = nullptr; }
    s->n_items = (int)items.size();
    FRMC_CUDA(cudaMalloc(&s->d_items, std::max<size_t>(blob.size(), 16)));
    FRMC_CUDA(cudaMemcpyAsync(s->d_items, blob.data(), blob.size(), cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    s->items_shard = shard; s->items_nshards = nshards;
    return FRMC_OK;
}

static int sync_models(frmc_store *s)
{
    if (!s->models_dirty) return FRMC_OK;
    size_t smem = 0;
    const size_t budget = 200 * 1024;
    for (auto &m : s->models) {
        const bool is_sq = (m.dev.kind == FRMC_KIND_SQ || m.dev.kind == FRMC_KIND_RSQ);
        const size_t base = sizeof(float) * ((size_t)(m.dev.hs + SQ_ROWS - 1) / SQ_ROWS * SQ_ROWS + (size_t)(m.dev.n_out + 31) / 32 * 32);
        FRMC_REQUIRE(base + (is_sq ? 2 * SQ_ROWS * 32 * sizeof(float) : 0) <= budget, FRMC_ELIMIT,
                     "model needs more than %zu B of shared memory in the epilogue", budget);
        m.dev.n_stages = 1;
        if (is_sq) {
            const int n_chunks = (m.dev.hs + SQ_ROWS - 1) / SQ_ROWS;
            int fit = (int)((budget - base) / (SQ_ROWS * 32 * sizeof(float)));
            m.dev.n_stages = std::max(2, std::min(std::min(fit, SQ_MAX_STAGES), n_chunks));
        }
        smem = std::max(smem, base + (is_sq ? (size_t)m.dev.n_stages * SQ_ROWS * 32 * sizeof(float) : 0));
    }
    FRMC_CUDA(cudaFuncSetAttribute(epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    // one shared-memory carveout for every kernel of the per-move pipeline: consecutive launches
    // with different carveouts make the SMs reconfigure (and drain) in between
    const int carve = cudaSharedmemCarveoutMaxShared;
    cudaFuncSetAttribute(epilogue_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(commit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(clear_delta_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(delta_kernel<MODE_IBC>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(delta_kernel<MODE_ORTHO_FAST>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(delta_kernel<MODE_TRI_FAST>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(delta_kernel<MODE_ORTHO_GEN>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(delta_kernel<MODE_TRI_GEN>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaGetLastError();
    s->epi_smem = smem;
    // fused path: epilogue CTA map, shared-memory attribute of every instantiation, co-residency check
    memset(&s->epi_map, 0, sizeof(s->epi_map));
    int n_epi = 0;
    bool fits = true;
    for (size_t mi = 0; mi < s->models.size(); ++mi) {
        const ModelDev &d = s->models[mi].dev;
        const bool is_sq = (d.kind == FRMC_KIND_SQ || d.kind == FRMC_KIND_RSQ);
        const int nblk = is_sq ? (d.n_out + 31) / 32 : 1;
        for (int x = 0; x < nblk; ++x) {
            if (n_epi >= 128 || x > 255) { fits = false; break; }
            s->epi_map.model[n_epi] = (unsigned char)mi; s->epi_map.slab[n_epi] = (unsigned char)x; ++n_epi;
        }
    }
    s->epi_map.n = fits ? n_epi : 0;
    s->fused_ok = false;
    if (s->use_fused && fits && n_epi >= 1 && n_epi <= s->ctx->sm_count) {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->dev);
        const int fsmem = (int)std::max<size_t>(smem, 48 * 1024);
        bool ok = coop != 0;
#define FUSED_ATTR(M) do { \
            if (cudaFuncSetAttribute(propose_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsmem) != cudaSuccess) ok = false; \
            cudaFuncSetAttribute(propose_kernel<M>, cudaFuncAttributePreferredSharedMemoryCarveout, carve); \
            int per_sm = 0; \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, propose_kernel<M>, EPI_THREADS, smem) != cudaSuccess || per_sm < 1) ok = false; \
        } while (0)
        FUSED_ATTR(MODE_IBC); FUSED_ATTR(MODE_ORTHO_FAST); FUSED_ATTR(MODE_TRI_FAST); FUSED_ATTR(MODE_ORTHO_GEN); FUSED_ATTR(MODE_TRI_GEN);
#undef FUSED_ATTR
        cudaGetLastError();
        s->fused_ok = ok;
    }
    // the persistent kernel keeps every S(Q) slab resident in shared memory for its whole life and carries the
    // model constants in its launch parameters: no ring refills, no per-evaluation refit schedule
    s->persist_ok = s->fused_ok && !s->models.empty();
    for (auto &m : s->models) {
        const bool is_sq = (m.dev.kind == FRMC_KIND_SQ || m.dev.kind == FRMC_KIND_RSQ);
        if (is_sq && (m.dev.hs + SQ_ROWS - 1) / SQ_ROWS > m.dev.n_stages) s->persist_ok = false;
        if (m.adjust_freq > 0) s->persist_ok = false;
    }
    if (s->persist_ok) {
        int per_sm = 0;
#define PERSIST_ATTR(M) do { \
            if (cudaFuncSetAttribute(propose_loop_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)) != cudaSuccess) s->persist_ok = false; \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, propose_loop_kernel<M>, EPI_THREADS, smem) != cudaSuccess || per_sm < 1) s->persist_ok = false; \
        } while (0)
        PERSIST_ATTR(MODE_IBC); PERSIST_ATTR(MODE_ORTHO_FAST); PERSIST_ATTR(MODE_TRI_FAST); PERSIST_ATTR(MODE_ORTHO_GEN); PERSIST_ATTR(MODE_TRI_GEN);
#undef PERSIST_ATTR
        cudaGetLastError();
    }
    // the batch kernel has the persistent kernel's requirements (resident slabs, constant refit flag)
    s->batch_ok = s->fused_ok && !s->models.empty();
    for (auto &m : s->models) {
        const bool is_sq = (m.dev.kind == FRMC_KIND_SQ || m.dev.kind == FRMC_KIND_RSQ);
        if (is_sq && (m.dev.hs + SQ_ROWS - 1) / SQ_ROWS > m.dev.n_stages) s->batch_ok = false;
    }
    if (s->batch_ok) {
        // S(Q) models without prior/window leave chi2 terms and every CTA sums them after the grid barrier
        s->batch_defer_mask = 0u;
        for (size_t mi = 0; mi < s->models.size(); ++mi) {
            const ModelDev &d = s->models[mi].dev;
            const bool is_sq = (d.kind == FRMC_KIND_SQ || d.kind == FRMC_KIND_RSQ);
            if (!is_sq || d.prior || d.window || d.pw_leaves > BATCH_DEFER_MAX_LEAVES || getenv("FRMC_BATCH_NO_DEFER")) continue;
            if (s->models[mi].adjust_freq > 0) continue;          // the refit is the last slab CTA's job (ticket path)
            s->batch_defer_mask |= 1u << mi;
        }
        int per_sm = 0;
#define BATCH_ATTR1(K) do { \
            if (cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)) != cudaSuccess) s->batch_ok = false; \
            cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, carve); \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, K, EPI_THREADS, smem) != cudaSuccess || per_sm < 1) s->batch_ok = false; \
        } while (0)
#define BATCH_ATTR(M) do { auto k0 = batch_kernel<M, false, false>; BATCH_ATTR1(k0); \
                           auto k2 = batch_kernel<M, false, true>; auto k3 = batch_kernel<M, true, true>; BATCH_ATTR1(k2); BATCH_ATTR1(k3); } while (0)
        BATCH_ATTR(MODE_IBC); BATCH_ATTR(MODE_ORTHO_FAST); BATCH_ATTR(MODE_TRI_FAST); BATCH_ATTR(MODE_ORTHO_GEN); BATCH_ATTR(MODE_TRI_GEN);
#undef BATCH_ATTR1
#undef BATCH_ATTR
        cudaGetLastError();
    }
    s->batch_ready = false;
    s->models_dirty = false;
    return FRMC_OK;
}

// per-launch copy of a model descriptor: the refit flag follows the engine's accepted count
static ModelDev launch_model(const frmc_store *s, const ModelHost &mh)
{
    ModelDev d = mh.dev;
    d.refit = (mh.adjust_freq > 0 && (s->accepted % (unsigned long long)mh.adjust_freq) == 0) ? 1 : 0;
    return d;
}

// totals + chi^2 of every model from the staged totals; results land in s->h_chi2 once h_seq == seq_expected
static int launch_epilogue_ms(frmc_store *s, const ModelSet &ms);

static int launch_epilogue(frmc_store *s)
{
    const int nm = (int)s->models.size();
    if (nm == 0) return FRMC_OK;
    int rc = sync_models(s);
    if (rc) return rc;
    ModelSet ms;
    memset(&ms, 0, sizeof(ms));
    ms.n = nm;
    for (int i = 0; i < nm; ++i) ms.m[i] = launch_model(s, s->models[i]);
    return launch_epilogue_ms(s, ms);
}

// the same launch with the model descriptors the caller prepared (the one-atom-fewer evaluation swaps constants)
static int launch_epilogue_ms(frmc_store *s, const ModelSet &ms)
{
    const int nm = ms.n;
    GridSet gs = make_gridset(s);
    int max_q = 1;
    for (int i = 0; i < nm; ++i)
        if (ms.m[i].kind == FRMC_KIND_SQ || ms.m[i].kind == FRMC_KIND_RSQ) max_q = std::max(max_q, ms.m[i].n_out);
    cudaEvent_t t0 = timing_begin(s);
    dim3 grid((unsigned)((max_q + 31) / 32), (unsigned)nm);
    epilogue_kernel<<<grid, EPI_THREADS, s->epi_smem, s->stream>>>(ms, gs, s->h_chi2, s->d_seq, s->h_seq,
                                                                   s->d_seq + FRMC_MAX_MODELS, s->d_stamps);
    FRMC_LAUNCH_CHECK();
    timing_end(s, TIME_EPILOGUE, t0);
    return FRMC_OK;
}

// wait for the epilogue of the launch numbered s->seq_expected: spin on the pinned per-model
// counters the epilogue publishes (a few hundred ns after the kernel retires), falling back
// to a stream synchronise to surface errors.
static int wait_epilogue(frmc_store *s)
{
    const size_t nm = s->models.size();
    if (nm == 0 || s->timing) { FRMC_CUDA(cudaStreamSynchronize(s->stream)); timing_flush(s); return FRMC_OK; }
    unsigned long long spins = 0;
    for (size_t m = 0; m < nm; ++m) {
        while (s->h_seq[m] != s->seq_expected) {
            _mm_pause();
            if (s->persist_running && (spins & 0xFFF) == 0xFFF && *reinterpret_cast<volatile unsigned int *>(&s->h_cmd->alive) == 0u) {
                // the kernel left on its watchdog just as the command was written: start a new one, send it again
                FRMC_CUDA(cudaStreamSynchronize(s->stream));
                s->persist_running = false;
                if (s->h_seq[m] == s->seq_expected) break;
                int rrc = launch_persistent(s, s->persist_mode);
                if (rrc) return rrc;
                --s->persist_cmds;                 // the same proposal, sent again
                write_cmd(s, CMD_EVAL, s->last_prev);
            }
            if ((++spins & 0xFFFFF) == 0) {           // every ~1M polls: make sure the stream is still alive
                cudaError_t e = cudaStreamQuery(s->stream);
                if (e == cudaSuccess) break;          // finished without publishing -> checked below
                if (e != cudaErrorNotReady) { set_error("stream error while waiting for chi2: %s", cudaGetErrorString(e)); return FRMC_ECUDA; }
            }
        }
        if (s->h_seq[m] != s->seq_expected) {
            FRMC_CUDA(cudaStreamSynchronize(s->stream));
            FRMC_REQUIRE(s->h_seq[m] == s->seq_expected, FRMC_ECUDA, "epilogue finished without publishing chi2 (seq %u, expected %u)",
                         (unsigned)s->h_seq[m], s->seq_expected);
        }
    }
    return FRMC_OK;
}

// numpy's FLOAT_pairwise_sum recursion for n terms, flattened: leaves left to right and the
// post-order combine steps (the left child's result lives in its first leaf's slot).
static int pairwise_walk(int o, int c, std::vector<int> &off, std::vector<int> &len, std::vector<int> &dst, std::vector<int> &src)
{
    if (c <= 128) { off.push_back(o); len.push_back(c); return (int)off.size() - 1; }
    int n2 = c / 2; n2 -= n2 % 8;
    int l = pairwise_walk(o, n2, off, len, dst, src);
    int r = pairwise_walk(o + n2, c - n2, off, len, dst, src);
    dst.push_back(l); src.push_back(r);
    return l;
}

static void pairwise_schedule(int n, std::vector<int> &sched, int &n_leaves)
{
    std::vector<int> off, len, dst, src;
    pairwise_walk(0, n, off, len, dst, src);
    n_leaves = (int)off.size();
    dst.resize(n_leaves, 0); src.resize(n_leaves, 0);
    sched.clear();
    sched.insert(sched.end(), off.begin(), off.end());
    sched.insert(sched.end(), len.begin(), len.end());
    sched.insert(sched.end(), dst.begin(), dst.end());
    sched.insert(sched.end(), src.begin(), src.end());
}

template <typename T>
static int dev_copy(ModelHost &mh, const T *src, size_t count, const T **dst)
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

// rD[p] = RN(1/D[p]) when div_by_const is proven equal to the IEEE division for pair p over the whole count range
// (validate_fastdiv_kernel on the device arrays d_w / d_D), NaN otherwise (the epilogue then divides)
static int validated_reciprocals(frmc_store *s, int n_pairs, const float *d_w, const float *d_D, const float *h_D, std::vector<float> &rD)
{
    int *d_ok = nullptr;
    float *d_rD = nullptr;
    std::vector<int> ok((size_t)n_pairs, 1);
    rD.assign((size_t)n_pairs, 0.f);
    FRMC_CUDA(cudaMalloc(&d_ok, sizeof(int) * n_pairs));
    FRMC_CUDA(cudaMalloc(&d_rD, sizeof(float) * n_pairs));
    FRMC_CUDA(cudaMemcpy(d_ok, ok.data(), sizeof(int) * n_pairs, cudaMemcpyHostToDevice));
    validate_fastdiv_kernel<<<dim3(64, (unsigned)n_pairs), 256, 0, s->stream>>>(d_w, d_D, n_pairs, d_ok, d_rD);
    ++g_launch_count;
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e == cudaSuccess) e = cudaMemcpy(ok.data(), d_ok, sizeof(int) * n_pairs, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(rD.data(), d_rD, sizeof(float) * n_pairs, cudaMemcpyDeviceToHost);
    cudaFree(d_ok); cudaFree(d_rD);
    if (e != cudaSuccess) { set_error("fast-division validation failed: %s", cudaGetErrorString(e)); return FRMC_ECUDA; }
    for (int p = 0; p < n_pairs; ++p) {
        const float Dp = h_D[p];
        const bool sane = (Dp == Dp) && !isinf(Dp) && fabsf(Dp) > 1e-30f && fabsf(Dp) < 1e30f && (rD[p] == rD[p]);
        if (!(ok[p] && sane)) rD[p] = __builtin_nanf("");
    }
    return FRMC_OK;
}

static int current_mode(frmc_store *s, const float *extra_lo, const float *extra_hi)
{
    float lo[3], hi[3];
    for (int c = 0; c < 3; ++c) {
        lo[c] = extra_lo ? std::min(s->lo[c], extra_lo[c]) : s->lo[c];
        hi[c] = extra_hi ? std::max(s->hi[c], extra_hi[c]) : s->hi[c];
    }
    if (s->force_general) hi[0] = INFINITY;
    return choose_mode_from_bounds(s->L.b, s->isPBC, lo, hi);
}

static int launch_cells_grid(frmc_store *s)
{
    long long cells = 0;
    for (auto &g : s->grids) cells = std::max(cells, 2 * g.dev.cells);
    return (int)std::max<long long>(1, std::min<long long>((cells + 255) / 256, (long long)s->ctx->sm_count * 2));
}

static TotalsCopy make_totals_copy(frmc_store *s)
{
    TotalsCopy tc;
    memset(&tc, 0, sizeof(tc));
    tc.n = (int)s->models.size();
    for (int m = 0; m < tc.n; ++m) { tc.len[m] = s->models[m].dev.n_out; tc.src[m] = s->models[m].dev.total; tc.dst[m] = s->models[m].total_committed; }
    return tc;
}

// apply a deferred accept / reject with the stand-alone kernels (needed before anything other than
// the next fused proposal looks at the device state)
static int flush_pending(frmc_store *s)
{
    { int prc = stop_persistent(s); if (prc) return prc; }      // anything that looks at the device state ends the run
    if (!s->pending) return FRMC_OK;
    GridSet gs = make_gridset(s);
    cudaEvent_t t0 = timing_begin(s);
    if (s->pending == 1) commit_kernel<<<launch_cells_grid(s), 256, 0, s->stream>>>(gs, s->d_atoms, s->d_prop, make_totals_copy(s));
    else clear_delta_kernel<<<launch_cells_grid(s), 256, 0, s->stream>>>(gs);
    FRMC_LAUNCH_CHECK();
    timing_end(s, TIME_COMMIT, t0);
    s->pending = 0;
    return FRMC_OK;
}

template <int MODE>
static int launch_fused_t(frmc_store *s)
{
    GridSet gs = make_gridset(s);
    ModelSet ms;
    memset(&ms, 0, sizeof(ms));
    ms.n = (int)s->models.size();
    for (int i = 0; i < ms.n; ++i) ms.m[i] = launch_model(s, s->models[i]);
    TotalsCopy tc = make_totals_copy(s);
    int npad = (int)s->npad, nEl = s->nEl, prev = s->pending;
    unsigned long long launch_no = ++s->fused_launches;
    if (prev) ++s->fused_resolves;
    unsigned long long resolve_no = s->fused_resolves;
    unsigned int *tickets = s->d_seq + FRMC_MAX_MODELS;
    volatile unsigned int *hseq = s->h_seq;
    void *args[] = {&s->d_atoms, &npad, &s->prop_in, &s->d_prop, &s->L, &gs, &nEl, &s->d_overflow, &ms, &s->epi_map, &prev, &tc,
                    &s->d_bars, &launch_no, &resolve_no, &s->h_chi2, &s->d_seq, &hseq, &tickets, &s->d_stamps};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)propose_kernel<MODE>, dim3((unsigned)s->ctx->sm_count), dim3(EPI_THREADS),
                                                args, s->epi_smem, s->stream);
    if (e != cudaSuccess) {
        --s->fused_launches; if (prev) --s->fused_resolves;
        set_error("cooperative launch of the fused propose kernel failed: %s", cudaGetErrorString(e));
        return FRMC_ECUDA;
    }
    ++g_launch_count;
    s->pending = 0;
    return FRMC_OK;
}

// ---- persistent per-move kernel: lifecycle
static void write_cmd(frmc_store *s, int op, int prev)
{
    HostCmd *h = s->h_cmd;
    const ProposalIn &in = s->prop_in;
    const int k = (op == CMD_EVAL) ? in.k : 0;
    for (int t = 1; t < k; ++t) {
        h->pos[t] = in.pos[t];
        h->moved[3 * t] = in.moved[3 * t]; h->moved[3 * t + 1] = in.moved[3 * t + 1]; h->moved[3 * t + 2] = in.moved[3 * t + 2];
    }
    _mm_sfence();                                     // entries 1.. are visible before the chunks that announce them
    const unsigned int seq = ++s->cmd_seq;
    unsigned int mx, my, mz;
    memcpy(&mx, &in.moved[0], 4); memcpy(&my, &in.moved[1], 4); memcpy(&mz, &in.moved[2], 4);
    const __m128i A = _mm_set_epi32((int)mx, in.pos[0], (int)((unsigned)op | ((unsigned)prev << 8) | ((unsigned)k << 16)), (int)seq);
    const __m128i B = _mm_set_epi32(0, (int)seq, (int)mz, (int)my);
    _mm_store_si128(reinterpret_cast<__m128i *>(&h->b_my0), B);      // one aligned 16-byte store per chunk
    _mm_store_si128(reinterpret_cast<__m128i *>(&h->a_seq), A);
    _mm_sfence();
    if (op == CMD_EVAL) ++s->persist_cmds;
}

static int stop_persistent(frmc_store *s)
{
    if (!s->persist_running) return FRMC_OK;
    if (*reinterpret_cast<volatile unsigned int *>(&s->h_cmd->alive)) write_cmd(s, CMD_QUIT, 0);
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    s->persist_running = false;
    return FRMC_OK;
}

namespace frmc {
int store_view(frmc_store *s, StoreView *out)
{
    FRMC_REQUIRE(s && out, FRMC_EINVAL, "NULL argument");
    FRMC_CUDA(cudaSetDevice(s->dev));
    int rc = stop_persistent(s);
    if (rc) return rc;
    out->dev = s->dev; out->stream = s->stream; out->sm_count = s->ctx->sm_count; out->ctx = s->ctx;
    out->atoms = s->d_atoms; out->orig = s->d_orig; out->n = s->n; out->npad = s->npad; out->inv = s->lay.inv.data();
    out->n0 = s->n0; out->layout_gen = s->layout_gen; out->rel2real = s->rel2real.empty() ? nullptr : s->rel2real.data();
    out->L = s->L; out->isPBC = s->isPBC;
    for (int c = 0; c < 3; ++c) { out->lo[c] = s->lo[c]; out->hi[c] = s->hi[c]; }
    out->pending = s->pending; out->prop = s->d_prop;
    return FRMC_OK;
}

int store_flush(frmc_store *s)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_CUDA(cudaSetDevice(s->dev));
    return flush_pending(s);
}
}  // namespace frmc

template <int MODE>
static int launch_persistent_t(frmc_store *s)
{
    GridSet gs = make_gridset(s);
    ModelSet ms;
    memset(&ms, 0, sizeof(ms));
    ms.n = (int)s->models.size();
    for (int i = 0; i < ms.n; ++i) ms.m[i] = launch_model(s, s->models[i]);
    TotalsCopy tc = make_totals_copy(s);
    int npad = (int)s->npad, nEl = s->nEl;
    unsigned int *tickets = s->d_seq + FRMC_MAX_MODELS;
    volatile unsigned int *hseq = s->h_seq;
    unsigned int first_seq = s->cmd_seq + 1;
    unsigned long long idle_ns = 2000000ull;          // 2 ms without a command: leave
    if (const char *e = getenv("FRMC_PERSIST_IDLE_US")) idle_ns = 1000ull * (unsigned long long)std::max(10, atoi(e));
    FRMC_CUDA(cudaMemsetAsync(s->d_pbars, 0, 2 * sizeof(unsigned long long), s->stream));
    FRMC_CUDA(cudaMemsetAsync(s->d_cmd, 0, 2 * sizeof(uint4), s->stream));            // relay chunks: 0 is never a command number
    *reinterpret_cast<volatile unsigned int *>(&s->h_cmd->alive) = 1u;
    _mm_sfence();
    void *args[] = {&s->d_atoms, &npad, &s->h_cmd, &s->d_cmd, &s->d_prop, &s->L, &gs, &nEl, &s->d_overflow, &ms, &s->epi_map, &tc,
                    &s->d_pbars, &first_seq, &idle_ns, &s->h_chi2, &s->d_seq, &hseq, &tickets};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)propose_loop_kernel<MODE>, dim3((unsigned)s->ctx->sm_count), dim3(EPI_THREADS),
                                                args, s->epi_smem, s->stream);
    if (e != cudaSuccess) {
        set_error("cooperative launch of the persistent propose kernel failed: %s", cudaGetErrorString(e));
        return FRMC_ECUDA;
    }
    ++g_launch_count;
    ++s->persist_launches;
    s->persist_running = true;
    return FRMC_OK;
}

static int launch_persistent(frmc_store *s, int mode)
{
    s->persist_mode = mode;
    switch (mode) {
        case MODE_IBC: return launch_persistent_t<MODE_IBC>(s);
        case MODE_ORTHO_FAST: return launch_persistent_t<MODE_ORTHO_FAST>(s);
        case MODE_TRI_FAST: return launch_persistent_t<MODE_TRI_FAST>(s);
        case MODE_ORTHO_GEN: return launch_persistent_t<MODE_ORTHO_GEN>(s);
        default: return launch_persistent_t<MODE_TRI_GEN>(s);
    }
}

// one proposal through the persistent kernel: (re)start it when needed, send the command
static int propose_persistent(frmc_store *s, int mode)
{
    int rc = FRMC_OK;
    const bool alive = s->persist_running && *reinterpret_cast<volatile unsigned int *>(&s->h_cmd->alive) != 0u;
    if (!alive || s->persist_mode != mode) {
        rc = stop_persistent(s);                      // also reaps a kernel that left on its own (watchdog)
        if (rc) return rc;
        rc = launch_persistent(s, mode);
        if (rc) return rc;
    }
    s->last_prev = s->pending;
    write_cmd(s, CMD_EVAL, s->pending);
    s->pending = 0;
    return FRMC_OK;
}

// the per-move pipeline.  Fused path: ONE cooperative launch (resolve previous + delta pass +
// epilogue).  Fallback (timing mode, FRMC_NO_FUSED=1, too many epilogue CTAs): delta pass
// (proposal by value) + fused epilogue, two launches, after the pending resolution.
static int launch_delta(frmc_store *s, int mode);

static int launch_propose(frmc_store *s, int mode)
{
    int rc = FRMC_OK;
    if (s->models_dirty && (rc = stop_persistent(s))) return rc;     // the running kernel carries the old constants
    rc = sync_models(s);
    if (rc) return rc;
    if (s->persist_enabled && s->persist_ok && !s->timing) return propose_persistent(s, mode);
    if ((rc = stop_persistent(s))) return rc;
    if (s->fused_ok && !s->timing) {
        switch (mode) {
            case MODE_IBC: return launch_fused_t<MODE_IBC>(s);
            case MODE_ORTHO_FAST: return launch_fused_t<MODE_ORTHO_FAST>(s);
            case MODE_TRI_FAST: return launch_fused_t<MODE_TRI_FAST>(s);
            case MODE_ORTHO_GEN: return launch_fused_t<MODE_ORTHO_GEN>(s);
            default: return launch_fused_t<MODE_TRI_GEN>(s);
        }
    }
    if ((rc = flush_pending(s))) return rc;
    if ((rc = launch_delta(s, mode))) return rc;
    return launch_epilogue(s);
}

// the stand-alone delta pass of the staged proposal (s->prop_in)
static int launch_delta(frmc_store *s, int mode)
{
    GridSet gs = make_gridset(s);
    long long want = (s->npad + 256 * DELTA_UNROLL - 1) / (256 * DELTA_UNROLL);
    long long cap = (long long)s->ctx->sm_count * 8;
    int grid = (int)std::max<long long>(1, std::min(want, cap));
#define LAUNCH_DELTA(M) delta_kernel<M><<<grid, 256, 0, s->stream>>>(s->d_atoms, (int)s->npad, s->prop_in, s->d_prop, s->L, gs, s->nEl, s->d_overflow, s->d_stamps)
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
    return FRMC_OK;
}

static int stage_proposal(frmc_store *s, const int32_t *indexes, int k, const float *moved)
{
    FRMC_REQUIRE(s && indexes && moved, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(k >= 1 && k <= FRMC_MAX_GROUP, FRMC_ELIMIT, "group size %d outside 1..%d", k, FRMC_MAX_GROUP);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is already staged; accept or reject it first");
    FRMC_REQUIRE(!s->grids.empty(), FRMC_ESTATE, "no grid registered");
    for (auto &g : s->grids) FRMC_REQUIRE(g.valid, FRMC_ESTATE, "call frmc_compute_data before proposing moves");
    ProposalIn &h = s->prop_in;
    h.k = k;
    for (int c = 0; c < 3; ++c) { s->prop_lo[c] = INFINITY; s->prop_hi[c] = -INFINITY; }
    bool finite = true;
    for (int t = 0; t < k; ++t) {
        FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < s->n, FRMC_EINVAL, "atom index %d outside 0..%lld", indexes[t], (long long)s->n - 1);
        h.pos[t] = pos_of(s, indexes[t]);
        for (int c = 0; c < 3; ++c) {
            float v = moved[3 * t + c];
            h.moved[3 * t + c] = v;
            if (!(v == v) || isinf(v)) finite = false;
            s->prop_lo[c] = std::min(s->prop_lo[c], v);
            s->prop_hi[c] = std::max(s->prop_hi[c], v);
        }
    }
    FRMC_REQUIRE(finite, FRMC_EINVAL, "moved coordinates contain NaN or Inf");
    return FRMC_OK;
}


// ---- runs of proposals resolved on the device
static int batch_prepare(frmc_store *s)
{
    if (s->batch_ready) return FRMC_OK;
    for (void *p : s->batch_owned) cudaFree(p);
    s->batch_owned.clear();
    BatchDev &bd = s->bdev;
    memset(&bd, 0, sizeof(bd));
    auto alloc = [&](void **out, size_t bytes) -> int {
        FRMC_CUDA(cudaMalloc(out, bytes));
        s->batch_owned.push_back(*out);
        FRMC_CUDA(cudaMemsetAsync(*out, 0, bytes, s->stream));
        return FRMC_OK;
    };
    int rc;
    for (size_t g = 0; g < s->grids.size(); ++g) {
        const GridDev &G = s->grids[g].dev;
        if ((rc = alloc((void **)&bd.bsym[g], sizeof(int) * ((size_t)BATCH_MAX_PROPS * G.nsym * G.g.hs + 4)))) return rc;
        if ((rc = alloc((void **)&bd.bdelta[g], sizeof(int) * ((size_t)BATCH_MAX_PROPS * 2 * G.cells + 4)))) return rc;
    }
    for (size_t m = 0; m < s->models.size(); ++m) {
        if ((rc = alloc((void **)&bd.btotal[m], sizeof(float) * (size_t)BATCH_MAX_PROPS * s->models[m].dev.n_out))) return rc;
        if ((rc = alloc((void **)&bd.bterm[m], sizeof(float) * (size_t)BATCH_MAX_PROPS * s->models[m].dev.n_out))) return rc;
        bd.total_committed[m] = s->models[m].total_committed;
    }
    if ((rc = alloc((void **)&bd.res, sizeof(float) * BATCH_MAX_PROPS * 2 * FRMC_MAX_MODELS))) return rc;
    if ((rc = alloc((void **)&bd.ptotal, sizeof(float) * BATCH_MAX_PROPS))) return rc;
    if ((rc = alloc((void **)&bd.tickets, sizeof(unsigned int) * BATCH_MAX_PROPS * (FRMC_MAX_MODELS + 1)))) return rc;
    if ((rc = alloc((void **)&bd.bov, sizeof(unsigned long long) * BATCH_MAX_PROPS))) return rc;
    if ((rc = alloc((void **)&bd.run, sizeof(BatchRun)))) return rc;
    if (!getenv("FRMC_BATCH_NO_FLY") &&
        (rc = alloc((void **)&bd.corr, sizeof(CorrEv) * (size_t)(BATCH_MAX_PROPS * (BATCH_MAX_PROPS - 1) / 2) * CORR_PER_PAIR))) return rc;
    if (!s->d_bbars) {
        FRMC_CUDA(cudaMalloc(&s->d_bbars, sizeof(unsigned long long) * 2));
        FRMC_CUDA(cudaMemsetAsync(s->d_bbars, 0, sizeof(unsigned long long) * 2, s->stream));
    }
    if (!s->h_brun) FRMC_CUDA(cudaHostAlloc((void **)&s->h_brun, sizeof(BatchRun), cudaHostAllocDefault));
    if (!s->bev0) { FRMC_CUDA(cudaEventCreate(&s->bev0)); FRMC_CUDA(cudaEventCreate(&s->bev1)); }
    bd.n_groups = std::max(1, std::min(BATCH_MAX_GROUPS, s->ctx->sm_count / std::max(1, s->epi_map.n)));
    if (const char *e = getenv("FRMC_BATCH_GROUPS")) bd.n_groups = std::max(1, std::min(bd.n_groups, atoi(e)));
    bd.defer_mask = s->batch_defer_mask;
    bd.debug = getenv("FRMC_BATCH_DEBUG") ? atoi(getenv("FRMC_BATCH_DEBUG")) : 0;
    {
        // On-the-fly pair corrections pay when proposals of one launch see each other: expected number of in-range pairs
        // among 32 atoms placed at random = 496 x (sphere of the widest window) / (volume of the system).
        double rmax2 = 0.0;
        for (auto &g : s->grids) rmax2 = std::max(rmax2, (double)g.dev.g.t2max);
        const double r = sqrt(std::max(rmax2, 0.0));
        double vol;
        if (s->isPBC) {
            const float *b = s->L.b;
            vol = fabs((double)b[0] * ((double)b[4] * b[8] - (double)b[5] * b[7]) - (double)b[1] * ((double)b[3] * b[8] - (double)b[5] * b[6]) +
                       (double)b[2] * ((double)b[3] * b[7] - (double)b[4] * b[6]));
        } else {
            vol = 1.0;
            for (int c = 0; c < 3; ++c) vol *= std::max(1e-3, (double)s->hi[c] - (double)s->lo[c]);
        }
        const double expected = 496.0 * (4.0 / 3.0) * 3.14159265358979 * r * r * r / std::max(vol, 1e-30);
        s->batch_fly = bd.corr != nullptr && !(expected < 4.0);
        if (const char *e = getenv("FRMC_BATCH_FLY")) s->batch_fly = bd.corr != nullptr && atoi(e) != 0;
    }
    s->batch_ready = true;
    return FRMC_OK;
}

template <int MODE>
static int launch_batch_t(frmc_store *s, const BatchIn *in_arr, int n_batches, bool generated)
{
    GridSet gs = make_gridset(s);
    ModelSet ms;
    memset(&ms, 0, sizeof(ms));
    ms.n = (int)s->models.size();
    for (int i = 0; i < ms.n; ++i) { ms.m[i] = s->models[i].dev; ms.m[i].refit = 0; }
    int npad = (int)s->npad, nEl = s->nEl;
    unsigned long long *ovf = s->d_overflow;
    if (!s->d_bstamps && getenv("FRMC_BATCH_STAMPS")) {
        FRMC_CUDA(cudaMalloc(&s->d_bstamps, sizeof(long long) * BATCH_STAMP_TOTAL));
    }
    if (s->d_bstamps) FRMC_CUDA(cudaMemsetAsync(s->d_bstamps, 0, sizeof(long long) * BATCH_STAMP_TOTAL, s->stream));
    // sub-block culling of the delta pass against the widest d^2 window of the grids (frmc_set_block_culling(0): off)
    {
        float amax = 0.f;
        for (int c = 0; c < 3; ++c) amax = std::max(amax, std::max(fabsf(s->lo[c]), fabsf(s->hi[c])));
        s->bdev.box_eps = 1e-6f * (1.0f + amax);
    }
    GridParams gw;
    memset(&gw, 0, sizeof(gw));
    gw.t2max = gs.t2hi;
    CullParams cp = make_cull(s->L, MODE, gw);
    if (g_no_cull) cp.enabled = 0;
    const BatchIn *in_dev = generated ? s->d_bin : nullptr;
    const GenOut *gen = generated ? s->d_gen : nullptr;
    float4 *real = (generated && s->isPBC) ? s->d_real : nullptr;
    void *args[] = {&s->d_atoms, &npad, (void *)&in_arr, &n_batches, &s->L, &gs, &nEl, &ms, &s->epi_map, &s->bdev, &cp, &s->d_bbars, &ovf, &s->d_bstamps,
                    &in_dev, &gen, &real};
    // (three variants per geometry mode, not four: generated runs always take the kernel with the on-the-fly corrections)
    const void *kern = generated ? (const void *)batch_kernel<MODE, true, true>
                                 : (s->batch_fly ? (const void *)batch_kernel<MODE, false, true> : (const void *)batch_kernel<MODE, false, false>);
    cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)s->ctx->sm_count), dim3(EPI_THREADS), args, s->epi_smem, s->stream);
    if (e != cudaSuccess) {
        set_error("cooperative launch of the batch kernel failed: %s", cudaGetErrorString(e));
        return FRMC_ECUDA;
    }
    ++g_launch_count;
    ++s->batch_launches;
    return FRMC_OK;
}

// in_arr: DEVICE array of n_batches batches of host proposals (NULL / 0 for a generated batch, which sits in s->d_bin)
static int launch_batch(frmc_store *s, int mode, const BatchIn *in_arr, int n_batches, bool generated = false)
{
    switch (mode) {
        case MODE_IBC: return launch_batch_t<MODE_IBC>(s, in_arr, n_batches, generated);
        case MODE_ORTHO_FAST: return launch_batch_t<MODE_ORTHO_FAST>(s, in_arr, n_batches, generated);
        case MODE_TRI_FAST: return launch_batch_t<MODE_TRI_FAST>(s, in_arr, n_batches, generated);
        case MODE_ORTHO_GEN: return launch_batch_t<MODE_ORTHO_GEN>(s, in_arr, n_batches, generated);
        default: return launch_batch_t<MODE_TRI_GEN>(s, in_arr, n_batches, generated);
    }
}

// the engine's decision (Engine.py:3317-3325) on the host, for the sequential fallback of frmc_run_batch
static float total_standard_error(const float *chi2, const float *var2, int nm)
{
    float term[FRMC_MAX_MODELS];
    for (int i = 0; i < nm; ++i) term[i] = chi2[i] / var2[i];
    if (nm == 8) return ((term[0] + term[1]) + (term[2] + term[3])) + ((term[4] + term[5]) + (term[6] + term[7]));
    float t = 0.0f;
    for (int i = 0; i < nm; ++i) t = t + term[i];
    return t;
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
    s->ctx = c; s->dev = dev; s->n = n; s->n0 = n; s->nEl = nEl; s->isPBC = isPBC ? 1 : 0;
    for (int i = 0; i < 9; ++i) s->L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    s->h_mol.assign(mol, mol + n);
    s->h_el.assign(el, el + n);
    memset(&s->prop_in, 0, sizeof(s->prop_in));
    s->step_timing = getenv("FRMC_STEP_TIMING") != nullptr;
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
    if (cudaMalloc(&s->d_overflow, 2 * sizeof(unsigned long long)) != cudaSuccess) return fail("alloc");
    cudaMemset(s->d_overflow, 0, 2 * sizeof(unsigned long long));
    if (cudaMalloc(&s->d_prop, sizeof(Proposal)) != cudaSuccess) return fail("alloc");
    cudaMemset(s->d_prop, 0, sizeof(Proposal));
    if (cudaHostAlloc(&s->h_chi2, sizeof(float) * 3 * FRMC_MAX_MODELS, cudaHostAllocMapped) != cudaSuccess) return fail("pinned alloc");
    s->h_seq = reinterpret_cast<volatile unsigned int *>(s->h_chi2 + FRMC_MAX_MODELS);
    for (int i = 0; i < FRMC_MAX_MODELS; ++i) s->h_seq[i] = 0u;
    if (cudaMalloc(&s->d_seq, sizeof(unsigned int) * 2 * FRMC_MAX_MODELS) != cudaSuccess) return fail("alloc");
    cudaMemset(s->d_seq, 0, sizeof(unsigned int) * 2 * FRMC_MAX_MODELS);
    if (cudaMalloc(&s->d_bars, sizeof(unsigned long long) * 4) != cudaSuccess) return fail("alloc");
    cudaMemset(s->d_bars, 0, sizeof(unsigned long long) * 4);
    { const char *e = getenv("FRMC_NO_FUSED"); s->use_fused = !(e && e[0] == '1'); }
    if (cudaHostAlloc((void **)&s->h_cmd, sizeof(HostCmd), cudaHostAllocMapped) != cudaSuccess) return fail("pinned alloc");
    memset(s->h_cmd, 0, sizeof(HostCmd));
    if (cudaMalloc((void **)&s->d_cmd, sizeof(DevCmd)) != cudaSuccess) return fail("alloc");
    cudaMemset(s->d_cmd, 0, sizeof(DevCmd));
    if (cudaMalloc((void **)&s->d_pbars, sizeof(unsigned long long) * 2) != cudaSuccess) return fail("alloc");
    { const char *e = getenv("FRMC_PERSISTENT"); s->persist_enabled = (e && e[0] == '1'); }
    if (cudaMalloc(&s->d_stamps, sizeof(long long) * 16 * FRMC_MAX_MODELS) != cudaSuccess) return fail("alloc");
    cudaMemset(s->d_stamps, 0, sizeof(long long) * 16 * FRMC_MAX_MODELS);
    { long long big = 0x7FFFFFFFFFFFFFFFll; cudaMemcpy(s->d_stamps + 120, &big, sizeof(big), cudaMemcpyHostToDevice); }
    for (int i = 0; i < FRMC_MAX_MODELS; ++i) { s->h_chi2[i] = 0.f; s->chi2_staged[i] = 0.f; s->chi2_committed[i] = 0.f; }
    return s;
}

void frmc_store_destroy(frmc_store *s)
{
    if (s && s->step_timing && s->n_calls)
        fprintf(stderr, "[step timing] %llu proposals: launch call %.2f us, wait for chi2 %.2f us (per proposal)\n", s->n_calls,
                s->t_launch / s->n_calls, s->t_wait / s->n_calls);
    if (!s) return;
    cudaSetDevice(s->dev);
    storedist_release(s);
    storecoord_release(s);
    if (s->h_cmd) stop_persistent(s);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto &m : s->models) {
        for (void *p : m.owned) cudaFree(p);
    }
    for (auto &g : s->grids) { cudaFree(g.dev.counts); cudaFree(g.dev.delta); cudaFree(g.dev.tot); cudaFree(g.dev.stot); }
    cudaFree(s->d_atoms); cudaFree(s->d_orig); cudaFree(s->d_items); cudaFree(s->d_next); cudaFree(s->d_mol);
    s->lists.release();
    cudaFree(s->d_overflow); cudaFree(s->d_bbox); cudaFree(s->d_prop); cudaFree(s->d_seq); cudaFree(s->d_stamps); cudaFree(s->d_bars);
    for (auto &p : s->ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : s->ev_pool) cudaEventDestroy(e);
    if (s->h_chi2) cudaFreeHost(s->h_chi2);
    if (s->h_cmd) cudaFreeHost(s->h_cmd);
    cudaFree(s->d_cmd); cudaFree(s->d_pbars);
    for (void *p : s->batch_owned) cudaFree(p);
    cudaFree(s->d_bstamps);
    cudaFree(s->d_real); cudaFree(s->d_inv); cudaFree(s->d_goff); cudaFree(s->d_gidx); cudaFree(s->d_bin); cudaFree(s->d_bins); cudaFree(s->d_gen); cudaFree(s->d_gout);
    cudaFree(s->d_bbars); cudaFree(s->d_brand); cudaFree(s->d_bout_chi2); cudaFree(s->d_bout_dec);
    if (s->h_brun) cudaFreeHost(s->h_brun);
    if (s->bev0) cudaEventDestroy(s->bev0);
    if (s->bev1) cudaEventDestroy(s->bev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

void *frmc_store_stream(frmc_store *s) { return s ? (void *)s->stream : nullptr; }

int frmc_store_set_coords(frmc_store *s, const float *coords, const float *basis)
{
    FRMC_REQUIRE(s && coords, FRMC_EINVAL, "NULL argument");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    if (basis) for (int i = 0; i < 9; ++i) s->L.b[i] = basis[i];
    if (!s->rel2real.empty()) {
        // atoms were removed since the layout was built: the caller's array has one row per remaining atom, so the
        // per-atom tables are compacted and the store starts a new "real" numbering
        std::vector<int32_t> mol((size_t)s->n), el((size_t)s->n);
        for (int64_t i = 0; i < s->n; ++i) { mol[i] = s->h_mol[s->rel2real[i]]; el[i] = s->h_el[s->rel2real[i]]; }
        s->h_mol.swap(mol); s->h_el.swap(el);
        s->rel2real.clear();
        s->n0 = s->n;
        cudaFree(s->d_mol); s->d_mol = nullptr;
    }
    int rc = upload_layout(s, coords);
    if (rc) return rc;
    cudaFree(s->d_inv); s->d_inv = nullptr;              // generated runs: new layout, new window, real coordinates again
    s->real_valid = false; s->force_general = false; s->gen_win_kind = 0;
    s->items_shard = s->items_nshards = -1;
    for (auto &g : s->grids) g.valid = false;
    s->state = 0;
    GridSet gs = make_gridset(s);
    if (gs.n) { clear_delta_kernel<<<launch_cells_grid(s), 256, 0, s->stream>>>(gs); FRMC_LAUNCH_CHECK(); }
    return FRMC_OK;
}

int frmc_store_move_atoms(frmc_store *s, const int32_t *indexes, int k, const float *moved)
{
    FRMC_REQUIRE(s && indexes && moved, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(k >= 1 && k <= FRMC_MAX_GROUP, FRMC_ELIMIT, "group size %d outside 1..%d", k, FRMC_MAX_GROUP);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    for (int t = 0; t < k; ++t) {
        FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < s->n, FRMC_EINVAL, "atom index %d outside 0..%lld", indexes[t], (long long)s->n - 1);
        for (int c = 0; c < 3; ++c) {
            const float v = moved[3 * t + c];
            FRMC_REQUIRE(v == v && !isinf(v), FRMC_EINVAL, "moved coordinates contain NaN or Inf");
        }
    }
    for (int t = 0; t < k; ++t) {
        // coordinates only: the record keeps its meta word (12-byte copy into the 16-byte record)
        FRMC_CUDA(cudaMemcpyAsync(reinterpret_cast<float *>(s->d_atoms + pos_of(s, indexes[t])), moved + 3 * t, sizeof(float) * 3,
                                  cudaMemcpyHostToDevice, s->stream));
        for (int c = 0; c < 3; ++c) { s->lo[c] = std::min(s->lo[c], moved[3 * t + c]); s->hi[c] = std::max(s->hi[c], moved[3 * t + c]); }
    }
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    s->real_valid = false;
    for (auto &g : s->grids) g.valid = false;          // running histograms (if any) no longer describe the coordinates
    return FRMC_OK;
}

int frmc_store_get_coords(frmc_store *s, float *coords_out)
{
    FRMC_REQUIRE(s && coords_out, FRMC_EINVAL, "NULL argument");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    std::vector<float> rec((size_t)s->npad * 4);
    FRMC_CUDA(cudaMemcpyAsync(rec.data(), s->d_atoms, sizeof(float4) * s->npad, cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < s->n; ++i) {
        int64_t p = pos_of(s, (int)i);
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
    { int frc = flush_pending(s); if (frc) return frc; }
    s->models_dirty = true;
    GridHost gh;
    memset(&gh.dev, 0, sizeof(gh.dev));
    gh.dev.g = make_grid(rmin, rmax, bin, hs);
    gh.dev.nsym = s->nEl * (s->nEl + 1) / 2;
    gh.dev.cells = (long long)s->nEl * s->nEl * hs;
    const long long ns = (long long)gh.dev.nsym * hs;
    FRMC_CUDA(cudaMalloc(&gh.dev.counts, sizeof(unsigned long long) * 2 * gh.dev.cells));
    FRMC_CUDA(cudaMalloc(&gh.dev.delta, sizeof(int) * 2 * gh.dev.cells));
    FRMC_CUDA(cudaMalloc(&gh.dev.tot, sizeof(int) * ns));
    FRMC_CUDA(cudaMalloc(&gh.dev.stot, sizeof(int) * ns));
    FRMC_CUDA(cudaMemsetAsync(gh.dev.counts, 0, sizeof(unsigned long long) * 2 * gh.dev.cells, s->stream));
    FRMC_CUDA(cudaMemsetAsync(gh.dev.delta, 0, sizeof(int) * 2 * gh.dev.cells, s->stream));
    FRMC_CUDA(cudaMemsetAsync(gh.dev.tot, 0, sizeof(int) * ns, s->stream));
    FRMC_CUDA(cudaMemsetAsync(gh.dev.stot, 0, sizeof(int) * ns, s->stream));
    s->grids.push_back(gh);
    return (int)s->grids.size() - 1;
}

int frmc_model_add(frmc_store *s, int grid, const frmc_model_desc *d)
{
    FRMC_REQUIRE(s && d, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(grid >= 0 && grid < (int)s->grids.size(), FRMC_EINVAL, "unknown grid %d", grid);
    FRMC_REQUIRE(s->models.size() < FRMC_MAX_MODELS, FRMC_ELIMIT, "at most %d models per store", FRMC_MAX_MODELS);
    FRMC_REQUIRE(d->kind >= FRMC_KIND_PDF && d->kind <= FRMC_KIND_RSQ, FRMC_EINVAL, "unknown model kind %d", d->kind);
    FRMC_REQUIRE(d->n_pairs >= 1 && d->n_pairs <= EPI_MAX_PAIRS && d->pair_a && d->pair_b && d->pair_w && d->pair_D,
                 FRMC_EINVAL, "bad pair table (n_pairs=%d)", d->n_pairs);
    FRMC_REQUIRE(d->shell_volumes && d->prefactor && d->experimental && d->n_out >= 1, FRMC_EINVAL, "missing model arrays");
    const int hs = s->grids[grid].dev.g.hs;
    const bool is_sq = (d->kind == FRMC_KIND_SQ || d->kind == FRMC_KIND_RSQ);
    FRMC_REQUIRE(is_sq ? (d->gr2sq != nullptr) : (d->n_out == hs), FRMC_EINVAL,
                 "model output length %d inconsistent with grid histSize %d", d->n_out, hs);
    FRMC_REQUIRE(d->n_out <= 64 * PW_MAX_LEAVES, FRMC_ELIMIT, "model output too long (%d)", d->n_out);
    std::vector<int> psym(d->n_pairs);
    for (int p = 0; p < d->n_pairs; ++p) {
        FRMC_REQUIRE(d->pair_a[p] >= 0 && d->pair_a[p] < s->nEl && d->pair_b[p] >= 0 && d->pair_b[p] < s->nEl,
                     FRMC_EINVAL, "pair %d references an element outside 0..%d", p, s->nEl - 1);
        psym[p] = sym_index(d->pair_a[p], d->pair_b[p], s->nEl);
    }
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    ModelHost mh;
    memset(&mh.dev, 0, sizeof(mh.dev));
    mh.dev.kind = d->kind; mh.dev.grid = grid; mh.dev.n_pairs = d->n_pairs; mh.dev.n_out = d->n_out;
    mh.dev.hs = hs; mh.dev.sq_exact = d->sq_exact; mh.dev.scale = d->scale;
    int rc;
    if ((rc = dev_copy(mh, psym.data(), (size_t)d->n_pairs, &mh.dev.psym))) return rc;
    if ((rc = dev_copy(mh, d->pair_w, (size_t)d->n_pairs, &mh.dev.w))) return rc;
    if ((rc = dev_copy(mh, d->pair_D, (size_t)d->n_pairs, &mh.dev.D))) return rc;
    {   // reciprocal table for the 3-op exact division, each pair validated on the device
        std::vector<float> rD;
        if ((rc = validated_reciprocals(s, d->n_pairs, mh.dev.w, mh.dev.D, d->pair_D, rD))) return rc;
        if ((rc = dev_copy(mh, (const float *)rD.data(), rD.size(), &mh.dev.rD))) return rc;
        if (d->n_pairs <= EPI_INLINE_PAIRS)
            for (int p = 0; p < d->n_pairs; ++p) {
                mh.dev.i_psym[p] = psym[p]; mh.dev.i_w[p] = d->pair_w[p]; mh.dev.i_D[p] = d->pair_D[p]; mh.dev.i_rD[p] = rD[p];
            }
    }
    if ((rc = dev_copy(mh, d->shell_volumes, (size_t)hs, &mh.dev.sv))) return rc;
    if ((rc = dev_copy(mh, d->prefactor, (size_t)hs, &mh.dev.pref))) return rc;
    if ((rc = dev_copy(mh, d->shape, (size_t)hs, &mh.dev.shape))) return rc;
    if ((rc = dev_copy(mh, d->experimental, (size_t)d->n_out, &mh.dev.expv))) return rc;
    if ((rc = dev_copy(mh, d->data_weights, (size_t)d->n_out, &mh.dev.wts))) return rc;
    if (is_sq) {
        // pre-tiled, zero-padded copy: [Q slab of 32][r/4][lane][r%4], rows padded to a multiple of SQ_ROWS,
        // so that one 64-row chunk of one slab is 8 KB contiguous (one TMA bulk copy) and four
        // consecutive rows of a column sit in one 16-byte word
        const int nslab = (d->n_out + 31) / 32;
        const int rows_pad = (hs + SQ_ROWS - 1) / SQ_ROWS * SQ_ROWS;
        mh.dev.nq_pad = nslab * 32;
        std::vector<float> tiled((size_t)nslab * rows_pad * 32, 0.0f);
        for (int r = 0; r < hs; ++r)
            for (int q = 0; q < d->n_out; ++q)
                tiled[(((size_t)(q / 32) * (rows_pad / 4) + r / 4) * 32 + (q % 32)) * 4 + (r % 4)] = d->gr2sq[(size_t)r * d->n_out + q];
        if ((rc = dev_copy(mh, (const float *)tiled.data(), tiled.size(), &mh.dev.gr2sq))) return rc;
    }
    std::vector<int> sched;
    pairwise_schedule(d->n_out, sched, mh.dev.pw_leaves);
    FRMC_REQUIRE(mh.dev.pw_leaves <= PW_MAX_LEAVES, FRMC_ELIMIT, "model output too long (%d)", d->n_out);
    if ((rc = dev_copy(mh, (const int *)sched.data(), sched.size(), &mh.dev.pw_sched))) return rc;
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

int frmc_model_set_shape(frmc_store *s, int model, const float *shape)
{
    FRMC_REQUIRE(s && model >= 0 && model < (int)s->models.size(), FRMC_EINVAL, "unknown model %d", model);
    ModelHost &mh = s->models[model];
    FRMC_REQUIRE(mh.dev.kind == FRMC_KIND_PDF || mh.dev.kind == FRMC_KIND_PCF, FRMC_EINVAL, "only r-space models carry a shape array");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }          // also ends a persistent run: its L1 may hold the old array
    if (!shape) {
        if (mh.dev.shape) { mh.dev.shape = nullptr; s->models_dirty = true; }
        return FRMC_OK;
    }
    if (!mh.dev.shape) {
        void *p = nullptr;
        FRMC_CUDA(cudaMalloc(&p, sizeof(float) * mh.dev.hs));
        mh.owned.push_back(p);
        mh.dev.shape = (const float *)p;
        s->models_dirty = true;
    }
    FRMC_CUDA(cudaMemcpyAsync((void *)mh.dev.shape, shape, sizeof(float) * mh.dev.hs, cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    return FRMC_OK;
}

// replace one optional per-model array (device copy owned by the model); n == 0 / NULL switches it off
static int set_model_array(frmc_store *s, ModelHost &mh, const float **slot, const float *src, size_t n)
{
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    s->models_dirty = true;                    // the descriptor travels by value with every launch
    if (!src || n == 0) { *slot = nullptr; return FRMC_OK; }
    void *p = nullptr;
    FRMC_CUDA(cudaMalloc(&p, sizeof(float) * n));
    mh.owned.push_back(p);
    FRMC_CUDA(cudaMemcpyAsync(p, src, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    *slot = (const float *)p;
    return FRMC_OK;
}

int frmc_model_set_window(frmc_store *s, int model, const float *window, int n_window)
{
    FRMC_REQUIRE(s && model >= 0 && model < (int)s->models.size(), FRMC_EINVAL, "unknown model %d", model);
    ModelHost &mh = s->models[model];
    FRMC_REQUIRE(n_window >= 0 && n_window <= mh.dev.n_out, FRMC_EINVAL, "window length %d outside 0..%d", n_window, mh.dev.n_out);
    mh.dev.n_window = window ? n_window : 0;
    return set_model_array(s, mh, &mh.dev.window, window, (size_t)n_window);
}

int frmc_model_set_multiframe_prior(frmc_store *s, int model, const float *prior, float weight)
{
    FRMC_REQUIRE(s && model >= 0 && model < (int)s->models.size(), FRMC_EINVAL, "unknown model %d", model);
    ModelHost &mh = s->models[model];
    mh.dev.mf_weight = weight;
    return set_model_array(s, mh, &mh.dev.prior, prior, (size_t)mh.dev.n_out);
}

int frmc_model_set_adjust(frmc_store *s, int model, int frequency, float sf_min, float sf_max)
{
    FRMC_REQUIRE(s && model >= 0 && model < (int)s->models.size(), FRMC_EINVAL, "unknown model %d", model);
    FRMC_REQUIRE(frequency >= 0, FRMC_EINVAL, "negative frequency");
    if ((s->models[model].adjust_freq > 0) != (frequency > 0)) s->models_dirty = true;   // which kernels take the model depends on it
    s->models[model].adjust_freq = frequency;
    s->models[model].dev.sf_min = sf_min;
    s->models[model].dev.sf_max = sf_max;
    return FRMC_OK;
}

int frmc_model_get_scale(frmc_store *s, int model, float *committed, float *last_used)
{
    FRMC_REQUIRE(s && model >= 0 && model < (int)s->models.size(), FRMC_EINVAL, "unknown model %d", model);
    if (committed) *committed = s->models[model].dev.scale;
    if (last_used) *last_used = s->models[model].sf_staged;
    return FRMC_OK;
}

int frmc_store_set_persistent(frmc_store *s, int on)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    if (!on) { int rc = stop_persistent(s); if (rc) return rc; }
    s->persist_enabled = on != 0;
    return FRMC_OK;
}

int frmc_store_persistent_stats(frmc_store *s, uint64_t *kernel_launches, uint64_t *commands)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    if (kernel_launches) *kernel_launches = s->persist_launches;
    if (commands) *commands = s->persist_cmds;
    return FRMC_OK;
}

int frmc_store_set_accepted(frmc_store *s, uint64_t accepted)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    s->accepted = accepted;
    return FRMC_OK;
}

int frmc_compute_data_shard(frmc_store *s, int shard, int nshards)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(nshards >= 1 && shard >= 0 && shard < nshards, FRMC_EINVAL, "bad shard %d of %d", shard, nshards);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_CUDA(cudaSetDevice(s->dev));
    int rc = flush_pending(s);
    if (rc) return rc;
    const int mode = current_mode(s, nullptr, nullptr);
    rc = upload_items(s, shard, nshards);
    if (rc) return rc;
    for (auto &g : s->grids) {
        FRMC_CUDA(cudaMemsetAsync(g.dev.counts, 0, sizeof(unsigned long long) * 2 * g.dev.cells, s->stream));
        FRMC_CUDA(cudaMemsetAsync(g.dev.delta, 0, sizeof(int) * 2 * g.dev.cells, s->stream));
        FRMC_CUDA(cudaMemsetAsync(s->d_next, 0, sizeof(int) * 4, s->stream));
        if (s->n_items > 0) {
            cudaEvent_t t0 = timing_begin(s);
            FRMC_CUDA(cudaMemsetAsync(s->d_overflow + 1, 0, sizeof(unsigned long long), s->stream));
            rc = full_hist_launch(s->stream, s->ctx->sm_count, mode, s->d_atoms, s->d_orig, s->npad, s->d_bbox, s->d_items,
                                  s->n_items, s->n_pairs, s->lists, s->d_mol, s->mol_span, s->L, g.dev.g, s->nEl, g.dev.counts, s->d_overflow);
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
    if (flush_pending(s)) return nullptr;
    if (n_cells) *n_cells = 2 * s->grids[grid].dev.cells;
    return s->grids[grid].dev.counts;
}

int frmc_finalize_data(frmc_store *s, float *chi2)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    FRMC_CUDA(cudaMemsetAsync(s->d_next + 2, 0, sizeof(int), s->stream));      // the too-big flag of symmetrise_kernel
    for (auto &g : s->grids) {
        const long long ns = (long long)g.dev.nsym * g.dev.g.hs;
        int grid = (int)std::max<long long>(1, std::min<long long>((ns + 255) / 256, (long long)s->ctx->sm_count * 2));
        symmetrise_kernel<<<grid, 256, 0, s->stream>>>(g.dev, s->nEl, s->d_next + 2);
        FRMC_LAUNCH_CHECK();
    }
    int too_big = 0;
    FRMC_CUDA(cudaMemcpyAsync(&too_big, s->d_next + 2, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    ++s->seq_expected;
    int rc = launch_epilogue(s);
    if (rc) return rc;
    for (auto &m : s->models)
        FRMC_CUDA(cudaMemcpyAsync(m.total_committed, m.dev.total, sizeof(float) * m.dev.n_out, cudaMemcpyDeviceToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    timing_flush(s);
    FRMC_REQUIRE(!too_big, FRMC_ELIMIT, "a symmetrised histogram cell exceeds 2^30 counts (int32 running totals)");
    for (size_t i = 0; i < s->models.size(); ++i) {
        s->chi2_committed[i] = s->h_chi2[i];
        s->models[i].sf_staged = s->h_chi2[2 * FRMC_MAX_MODELS + i];
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
    int rc = stage_proposal(s, indexes, k, moved);
    if (rc) return rc;
    FRMC_CUDA(cudaSetDevice(s->dev));
    const int mode = current_mode(s, s->prop_lo, s->prop_hi);
    ++s->seq_expected;
    std::chrono::steady_clock::time_point c0, c1, c2;
    if (s->step_timing) c0 = std::chrono::steady_clock::now();
    rc = launch_propose(s, mode);
    if (rc) return rc;
    if (s->step_timing) c1 = std::chrono::steady_clock::now();
    rc = wait_epilogue(s);
    if (rc) return rc;
    if (s->step_timing) {
        c2 = std::chrono::steady_clock::now();
        s->t_launch += std::chrono::duration<double, std::micro>(c1 - c0).count();
        s->t_wait += std::chrono::duration<double, std::micro>(c2 - c1).count();
        ++s->n_calls;
    }
    for (size_t i = 0; i < s->models.size(); ++i) {
        s->chi2_staged[i] = s->h_chi2[i];
        s->models[i].sf_staged = s->h_chi2[2 * FRMC_MAX_MODELS + i];
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
    // the device-side commit is deferred: the next fused proposal resolves it in its first phase,
    // anything else that looks at the device state calls flush_pending() first
    s->pending = 1;
    s->real_valid = false;
    for (int c = 0; c < 3; ++c) { s->lo[c] = std::min(s->lo[c], s->prop_lo[c]); s->hi[c] = std::max(s->hi[c], s->prop_hi[c]); }
    for (size_t i = 0; i < s->models.size(); ++i) {
        s->chi2_committed[i] = s->chi2_staged[i];
        // accept_move: _set_fitted_scale_factor_value(self._fittedScaleFactor) (PairDistributionConstraints.py:1150)
        if (s->models[i].adjust_freq > 0) s->models[i].dev.scale = s->models[i].sf_staged;
    }
    ++s->accepted;
    s->state = 0;
    if (!(s->fused_ok && !s->timing)) return flush_pending(s);
    return FRMC_OK;
}

int frmc_reject(frmc_store *s)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->state == 1, FRMC_ESTATE, "no staged proposal to reject");
    FRMC_CUDA(cudaSetDevice(s->dev));
    s->pending = 2;
    s->state = 0;
    if (!(s->fused_ok && !s->timing)) return flush_pending(s);
    return FRMC_OK;
}

int frmc_step(frmc_store *s, int previous, const int32_t *indexes, int k, const float *moved, float *chi2_after)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    if (s->state == 1) {
        FRMC_REQUIRE(previous == 0 || previous == 1, FRMC_EINVAL, "a proposal is staged: previous must be 1 (accept) or 0 (reject)");
        int rc = previous ? frmc_accept(s) : frmc_reject(s);
        if (rc) return rc;
    }
    return frmc_propose(s, indexes, k, moved, chi2_after);
}

// ---- dynamic N and persisted state ---------------------------------------------------
int64_t frmc_store_n_atoms(frmc_store *s) { return s ? s->n : -1; }

int frmc_model_set_constants(frmc_store *s, int model, const float *pair_w, const float *pair_D, const float *prefactor)
{
    FRMC_REQUIRE(s && model >= 0 && model < (int)s->models.size(), FRMC_EINVAL, "unknown model %d", model);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }          // also ends a persistent run (descriptors travel by value)
    ModelHost &mh = s->models[model];
    const int np = mh.dev.n_pairs;
    if (pair_w) FRMC_CUDA(cudaMemcpyAsync((void *)mh.dev.w, pair_w, sizeof(float) * np, cudaMemcpyHostToDevice, s->stream));
    if (pair_D) FRMC_CUDA(cudaMemcpyAsync((void *)mh.dev.D, pair_D, sizeof(float) * np, cudaMemcpyHostToDevice, s->stream));
    if (prefactor) FRMC_CUDA(cudaMemcpyAsync((void *)mh.dev.pref, prefactor, sizeof(float) * mh.dev.hs, cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    if (pair_w || pair_D) {
        std::vector<float> hD((size_t)np), hw((size_t)np), rD;
        FRMC_CUDA(cudaMemcpy(hD.data(), mh.dev.D, sizeof(float) * np, cudaMemcpyDeviceToHost));
        FRMC_CUDA(cudaMemcpy(hw.data(), mh.dev.w, sizeof(float) * np, cudaMemcpyDeviceToHost));
        int rc = validated_reciprocals(s, np, mh.dev.w, mh.dev.D, hD.data(), rD);
        if (rc) return rc;
        FRMC_CUDA(cudaMemcpy((void *)mh.dev.rD, rD.data(), sizeof(float) * np, cudaMemcpyHostToDevice));
        if (np <= EPI_INLINE_PAIRS)
            for (int p = 0; p < np; ++p) { mh.dev.i_w[p] = hw[p]; mh.dev.i_D[p] = hD[p]; mh.dev.i_rD[p] = rD[p]; }
    }
    s->models_dirty = true;
    return FRMC_OK;
}

int frmc_propose_amputation(frmc_store *s, int32_t index, const frmc_amputation_desc *descs, int allow_fit, float *chi2_out)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_REQUIRE(!s->grids.empty(), FRMC_ESTATE, "no grid registered");
    for (auto &g : s->grids) FRMC_REQUIRE(g.valid, FRMC_ESTATE, "call frmc_compute_data before removing atoms");
    FRMC_REQUIRE(index >= 0 && index < s->n, FRMC_EINVAL, "atom index %d outside 0..%lld", index, (long long)s->n - 1);
    FRMC_REQUIRE(s->n >= 2, FRMC_ELIMIT, "cannot remove the last atom");
    FRMC_CUDA(cudaSetDevice(s->dev));
    int rc = flush_pending(s);
    if (rc) return rc;
    if ((rc = sync_models(s))) return rc;
    // the atom's row leaves the histograms: a proposal whose "moved" position pairs with nothing (NaN fails every
    // range test), i.e. data - compute_before_move's activeAtomsDataBeforeMove (PairDistributionConstraints.py:1181-1184)
    ProposalIn &h = s->prop_in;
    h.k = 1;
    h.pos[0] = pos_of(s, index);
    h.moved[0] = h.moved[1] = h.moved[2] = __builtin_nanf("");
    for (int c = 0; c < 3; ++c) { s->prop_lo[c] = s->lo[c]; s->prop_hi[c] = s->hi[c]; }
    const int mode = current_mode(s, nullptr, nullptr);
    ++s->seq_expected;
    if ((rc = launch_delta(s, mode))) return rc;
    const int nm = (int)s->models.size();
    ModelSet ms;
    memset(&ms, 0, sizeof(ms));
    ms.n = nm;
    for (int i = 0; i < nm; ++i) {
        ModelHost &mh = s->models[i];
        ModelDev d = launch_model(s, mh);
        if (!allow_fit) d.refit = 0;             // _set_adjust_scale_factor_frequency(0) around the evaluation (:1195-1197)
        const frmc_amputation_desc *a = descs ? descs + i : nullptr;
        if (a && (a->pair_w || a->pair_D)) {
            FRMC_REQUIRE(a->pair_w && a->pair_D, FRMC_EINVAL, "model %d: pair_w and pair_D come together", i);
            if (!mh.amp_w) {
                void *p = nullptr;
                FRMC_CUDA(cudaMalloc(&p, sizeof(float) * 3 * d.n_pairs)); mh.owned.push_back(p);
                mh.amp_w = (float *)p; mh.amp_D = mh.amp_w + d.n_pairs; mh.amp_rD = mh.amp_D + d.n_pairs;
            }
            std::vector<float> nan((size_t)d.n_pairs, __builtin_nanf(""));     // IEEE division for this one evaluation
            FRMC_CUDA(cudaMemcpyAsync(mh.amp_w, a->pair_w, sizeof(float) * d.n_pairs, cudaMemcpyHostToDevice, s->stream));
            FRMC_CUDA(cudaMemcpyAsync(mh.amp_D, a->pair_D, sizeof(float) * d.n_pairs, cudaMemcpyHostToDevice, s->stream));
            FRMC_CUDA(cudaMemcpyAsync(mh.amp_rD, nan.data(), sizeof(float) * d.n_pairs, cudaMemcpyHostToDevice, s->stream));
            d.w = mh.amp_w; d.D = mh.amp_D; d.rD = mh.amp_rD;
            if (d.n_pairs <= EPI_INLINE_PAIRS)
                for (int p = 0; p < d.n_pairs; ++p) { d.i_w[p] = a->pair_w[p]; d.i_D[p] = a->pair_D[p]; d.i_rD[p] = nan[p]; }
        }
        if (a && a->prefactor) {
            if (!mh.amp_pref) {
                void *p = nullptr;
                FRMC_CUDA(cudaMalloc(&p, sizeof(float) * d.hs)); mh.owned.push_back(p);
                mh.amp_pref = (float *)p;
            }
            FRMC_CUDA(cudaMemcpyAsync(mh.amp_pref, a->prefactor, sizeof(float) * d.hs, cudaMemcpyHostToDevice, s->stream));
            d.pref = mh.amp_pref;
        }
        ms.m[i] = d;
    }
    if (nm) {
        if ((rc = launch_epilogue_ms(s, ms))) return rc;
        if ((rc = wait_epilogue(s))) return rc;
    } else {
        FRMC_CUDA(cudaStreamSynchronize(s->stream));
    }
    for (int i = 0; i < nm; ++i) {
        s->chi2_staged[i] = s->h_chi2[i];
        s->models[i].sf_staged = s->h_chi2[2 * FRMC_MAX_MODELS + i];
        if (chi2_out) chi2_out[i] = s->h_chi2[i];
    }
    s->amp_rel = index;
    s->state = 2;
    return FRMC_OK;
}

int frmc_accept_amputation(frmc_store *s)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->state == 2, FRMC_ESTATE, "no staged amputation to accept");
    FRMC_CUDA(cudaSetDevice(s->dev));
    s->pending = 1;
    int rc = flush_pending(s);                 // counts += delta, totals, model totals
    if (rc) return rc;
    // the record becomes padding: NaN coordinates pair with nothing, PAD_META keeps it out of every sweep
    const int pos = s->prop_in.pos[0];
    const float nanv = __builtin_nanf("");
    float rec[4] = {nanv, nanv, nanv, 0.f};
    const uint32_t pad = PAD_META, none = 0xFFFFFFFFu;
    memcpy(&rec[3], &pad, 4);
    FRMC_CUDA(cudaMemcpyAsync(s->d_atoms + pos, rec, sizeof(float4), cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaMemcpyAsync(s->d_orig + pos, &none, sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    if (s->rel2real.empty()) {
        s->rel2real.resize((size_t)s->n);
        for (int64_t i = 0; i < s->n; ++i) s->rel2real[i] = (int32_t)i;
    }
    s->rel2real.erase(s->rel2real.begin() + s->amp_rel);
    --s->n;
    for (size_t i = 0; i < s->models.size(); ++i) {
        s->chi2_committed[i] = s->chi2_staged[i];
        s->models[i].dev.scale = s->models[i].sf_staged;      // accept_amputation: _set_fitted_scale_factor_value (:1227)
    }
    s->models_dirty = true;
    ++s->accepted;                              // Engine.__on_runtime_step_try_remove counts it as accepted (Engine.py:3259)
    s->amp_rel = -1;
    s->state = 0;
    return FRMC_OK;
}

int frmc_reject_amputation(frmc_store *s)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->state == 2, FRMC_ESTATE, "no staged amputation to reject");
    FRMC_CUDA(cudaSetDevice(s->dev));
    s->pending = 2;
    s->amp_rel = -1;
    s->state = 0;
    return flush_pending(s);
}

int frmc_import_data(frmc_store *s, int grid, const float *hintra, const float *hinter)
{
    FRMC_REQUIRE(s && hintra && hinter, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(grid >= 0 && grid < (int)s->grids.size(), FRMC_EINVAL, "unknown grid %d", grid);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    GridDev &G = s->grids[grid].dev;
    std::vector<long long> cnt((size_t)(2 * G.cells));
    for (int part = 0; part < 2; ++part) {
        const float *src = part ? hinter : hintra;
        for (long long c = 0; c < G.cells; ++c) {
            const float v = src[c];
            const long long iv = (long long)v;
            FRMC_REQUIRE(v == v && fabsf(v) < 9.0e18f && (float)iv == v, FRMC_EINVAL,
                         "saved histogram cell %lld of %s is not an integer count (%g)", c, part ? "inter" : "intra", (double)v);
            cnt[(size_t)part * G.cells + c] = iv;
        }
    }
    FRMC_CUDA(cudaMemcpyAsync(G.counts, cnt.data(), sizeof(long long) * cnt.size(), cudaMemcpyHostToDevice, s->stream));
    FRMC_CUDA(cudaMemsetAsync(G.delta, 0, sizeof(int) * 2 * G.cells, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    s->grids[grid].valid = true;
    return FRMC_OK;
}

int frmc_run_batch(frmc_store *s, int n, const int32_t *group_sizes, const int32_t *indexes, const float *moved,
                   const float *variance_sq, float tolerance, const float *rand, float *total_io,
                   float *chi2_out, int32_t *decisions, int32_t *n_rand_used, double *device_ms)
{
    FRMC_REQUIRE(s && indexes && moved && total_io, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(n >= 1, FRMC_EINVAL, "need at least one proposal");
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_REQUIRE(!s->grids.empty() && !s->models.empty(), FRMC_ESTATE, "a run of proposals needs at least one grid and one model");
    for (auto &g : s->grids) FRMC_REQUIRE(g.valid, FRMC_ESTATE, "call frmc_compute_data before proposing moves");
    const int nm = (int)s->models.size();
    float var2[FRMC_MAX_MODELS];
    for (int i = 0; i < FRMC_MAX_MODELS; ++i) var2[i] = (variance_sq && i < nm) ? variance_sq[i] : 1.0f;
    // offsets of the groups, sizes checked
    std::vector<int> first((size_t)n + 1, 0);
    for (int j = 0; j < n; ++j) {
        const int k = group_sizes ? group_sizes[j] : 1;
        FRMC_REQUIRE(k >= 1 && k <= FRMC_MAX_GROUP, FRMC_ELIMIT, "group size %d of proposal %d outside 1..%d", k, j, FRMC_MAX_GROUP);
        first[j + 1] = first[j] + k;
    }
    const int n_atoms = first[n];
    float plo[3] = {INFINITY, INFINITY, INFINITY}, phi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int t = 0; t < n_atoms; ++t) {
        FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < s->n, FRMC_EINVAL, "atom index %d outside 0..%lld", indexes[t], (long long)s->n - 1);
        for (int c = 0; c < 3; ++c) {
            const float v = moved[3 * t + c];
            FRMC_REQUIRE(v == v && !isinf(v), FRMC_EINVAL, "moved coordinates contain NaN or Inf");
            plo[c] = std::min(plo[c], v); phi[c] = std::max(phi[c], v);
        }
    }
    FRMC_CUDA(cudaSetDevice(s->dev));
    int rc = flush_pending(s);
    if (rc) return rc;
    if ((rc = sync_models(s))) return rc;
    int used_rand = 0;
    if (device_ms) *device_ms = 0.0;
    if (!s->batch_ok || s->timing) {
        // sequential device path (one launch per proposal), same rule, same outputs: models the batch kernel
        // does not take (S(Q) slab larger than shared memory, scale-factor refit schedule)
        float total = *total_io;
        std::vector<float> chi2((size_t)FRMC_MAX_MODELS, 0.f);
        for (int j = 0; j < n; ++j) {
            rc = frmc_propose(s, indexes + first[j], first[j + 1] - first[j], moved + 3 * (size_t)first[j], chi2.data());
            if (rc) return rc;
            const float nt = total_standard_error(chi2.data(), var2, nm);
            int dec = 1;
            if (nt > total) {
                FRMC_REQUIRE(rand != nullptr, FRMC_EINVAL, "a worse proposal needs a random number and rand is NULL");
                dec = (rand[used_rand++] > tolerance) ? 0 : 2;
            }
            if (chi2_out) for (int m = 0; m < nm; ++m) chi2_out[(size_t)j * nm + m] = chi2[m];
            if (decisions) decisions[j] = dec;
            rc = dec ? frmc_accept(s) : frmc_reject(s);
            if (rc) return rc;
            if (dec) total = nt;
        }
        *total_io = total;
        if (n_rand_used) *n_rand_used = used_rand;
        return FRMC_OK;
    }
    FRMC_REQUIRE(rand != nullptr, FRMC_EINVAL, "rand is NULL (one pre-drawn number per proposal)");
    if ((rc = batch_prepare(s))) return rc;
    // conservative coordinate bounds for the wrap mode: every proposed position may become a stored one
    for (int c = 0; c < 3; ++c) { s->lo[c] = std::min(s->lo[c], plo[c]); s->hi[c] = std::max(s->hi[c], phi[c]); }
    const int mode = current_mode(s, nullptr, nullptr);
    // call-wide device arrays
    if (s->brand_cap < (size_t)n + 2 * BATCH_MAX_GROUPS) {
        cudaFree(s->d_brand); s->d_brand = nullptr; s->brand_cap = 0;
        FRMC_CUDA(cudaMalloc(&s->d_brand, sizeof(float) * ((size_t)n + 2 * BATCH_MAX_GROUPS)));
        s->brand_cap = (size_t)n + 2 * BATCH_MAX_GROUPS;
    }
    if (s->bout_cap < (size_t)n) {
        cudaFree(s->d_bout_chi2); cudaFree(s->d_bout_dec); s->d_bout_chi2 = nullptr; s->d_bout_dec = nullptr; s->bout_cap = 0;
        FRMC_CUDA(cudaMalloc(&s->d_bout_chi2, sizeof(float) * (size_t)n * FRMC_MAX_MODELS));
        FRMC_CUDA(cudaMalloc(&s->d_bout_dec, sizeof(int) * (size_t)n));
        s->bout_cap = (size_t)n;
    }
    FRMC_CUDA(cudaMemsetAsync(s->d_brand, 0, sizeof(float) * ((size_t)n + 2 * BATCH_MAX_GROUPS), s->stream));
    FRMC_CUDA(cudaMemcpyAsync(s->d_brand, rand, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    BatchDev &bd = s->bdev;
    bd.rand = s->d_brand; bd.out_chi2 = s->d_bout_chi2; bd.out_dec = s->d_bout_dec;
    for (int i = 0; i < FRMC_MAX_MODELS; ++i) bd.var2[i] = var2[i];
    bd.tol = tolerance;
    bd.rand_per_proposal = 0;
    for (int i = 0; i < FRMC_MAX_MODELS; ++i) bd.freq[i] = (i < nm) ? s->models[i].adjust_freq : 0;
    bd.accepted_base = s->accepted;
    s->real_valid = false;                      // accepted moves of this run bypass the real-coordinate array
    BatchRun run0;
    memset(&run0, 0, sizeof(run0));
    run0.total = *total_io;
    for (int m = 0; m < nm; ++m) { run0.cchi2[m] = s->chi2_committed[m]; run0.csf[m] = s->models[m].dev.scale; }
    *s->h_brun = run0;
    FRMC_CUDA(cudaMemcpyAsync(bd.run, s->h_brun, sizeof(BatchRun), cudaMemcpyHostToDevice, s->stream));
    if (device_ms) FRMC_CUDA(cudaEventRecord(s->bev0, s->stream));
    int done = 0;
    while (done < n) {
        // cut [done, n) into launches of at most BATCH_MAX_PROPS proposals / FRMC_MAX_GROUP atoms
        int j0 = done;
        std::vector<BatchIn> &batches = s->h_bins;
        batches.clear();
        while (j0 < n) {
            BatchIn in;
            memset(&in, 0, sizeof(in));
            in.out_base = j0;
            int np = 0, na = 0;
            while (j0 + np < n && np < BATCH_MAX_PROPS && na + (first[j0 + np + 1] - first[j0 + np]) <= FRMC_MAX_GROUP) {
                const int j = j0 + np;
                in.first[np] = na;
                unsigned int share = 0u;
                for (int t = first[j]; t < first[j + 1]; ++t) {
                    const int pos = pos_of(s, indexes[t]);
                    for (int v = 0; v < na; ++v)
                        if (in.pos[v] == pos) {
                            int jp = 0;
                            while (in.first[jp + 1] <= v && jp + 1 < np) ++jp;
                            share |= 1u << jp;
                        }
                }
                for (int t = first[j]; t < first[j + 1]; ++t, ++na) {
                    in.pos[na] = pos_of(s, indexes[t]);
                    in.moved[3 * na] = moved[3 * (size_t)t]; in.moved[3 * na + 1] = moved[3 * (size_t)t + 1]; in.moved[3 * na + 2] = moved[3 * (size_t)t + 2];
                }
                in.share[np] = share;
                ++np;
                in.first[np] = na;
            }
            in.n_prop = np; in.n_atoms = na;
            batches.push_back(in);
            j0 += np;
        }
        // all batches of the run go up at once; ONE launch works through up to BATCH_PER_LAUNCH of them (bounded so that a
        // kernel stays in the tens of milliseconds), later launches continue where it stops
        if (s->bins_cap < batches.size()) {
            FRMC_CUDA(cudaStreamSynchronize(s->stream));
            cudaFree(s->d_bins); s->d_bins = nullptr; s->bins_cap = 0;
            FRMC_CUDA(cudaMalloc(&s->d_bins, sizeof(BatchIn) * batches.size()));
            s->bins_cap = batches.size();
        }
        FRMC_CUDA(cudaMemcpyAsync(s->d_bins, batches.data(), sizeof(BatchIn) * batches.size(), cudaMemcpyHostToDevice, s->stream));
        for (size_t b0 = 0; b0 < batches.size(); b0 += BATCH_PER_LAUNCH) {
            const int nb = (int)std::min<size_t>(BATCH_PER_LAUNCH, batches.size() - b0);
            if ((rc = launch_batch(s, mode, s->d_bins + b0, nb))) return rc;
        }
        FRMC_CUDA(cudaMemcpyAsync(s->h_brun, bd.run, sizeof(BatchRun), cudaMemcpyDeviceToHost, s->stream));
        FRMC_CUDA(cudaStreamSynchronize(s->stream));
        const BatchRun &r = *s->h_brun;
        FRMC_REQUIRE(r.n_done > done && r.n_done <= n, FRMC_ECUDA, "batch kernel made no progress (done %d -> %d of %d)", done, r.n_done, n);
        done = r.n_done;
        if (done < n) {
            FRMC_REQUIRE(r.stopped, FRMC_ECUDA, "batch kernel ended early without a conflict (done %d of %d)", done, n);
            // a proposal moves an atom that an accepted proposal of its launch had moved: start a new launch at it
            s->h_brun->stopped = 0;
            FRMC_CUDA(cudaMemcpyAsync(bd.run, s->h_brun, sizeof(BatchRun), cudaMemcpyHostToDevice, s->stream));
        }
    }
    if (device_ms) {
        FRMC_CUDA(cudaEventRecord(s->bev1, s->stream));
        FRMC_CUDA(cudaEventSynchronize(s->bev1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->bev0, s->bev1);
        *device_ms = ms;
    }
    if (chi2_out) {
        FRMC_CUDA(cudaMemcpyAsync(chi2_out, s->d_bout_chi2, sizeof(float) * (size_t)n * nm, cudaMemcpyDeviceToHost, s->stream));
    }
    if (decisions) FRMC_CUDA(cudaMemcpyAsync(decisions, s->d_bout_dec, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    const BatchRun &r = *s->h_brun;
    *total_io = r.total;
    if (n_rand_used) *n_rand_used = r.n_rand;
    if (r.n_accepted > 0)
        for (int m = 0; m < nm; ++m) {
            s->chi2_committed[m] = r.cchi2[m]; s->chi2_staged[m] = r.cchi2[m];
            if (s->models[m].adjust_freq > 0) { s->models[m].dev.scale = r.csf[m]; s->models[m].sf_staged = r.csf[m]; }
        }
    s->accepted += (unsigned long long)r.n_accepted;
    s->batch_rounds += (unsigned long long)r.rounds;
    s->batch_proposals += (unsigned long long)n;
    return FRMC_OK;
}

// ---- device-generated runs of moves ---------------------------------------------------
int frmc_store_set_real_coords(frmc_store *s, const float *real, const float *rbasis)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    FRMC_REQUIRE(s->rel2real.empty(), FRMC_ESTATE, "atoms were removed from this store: re-create it before generated runs");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    if (!s->d_inv) {
        FRMC_CUDA(cudaMalloc(&s->d_inv, sizeof(int) * (size_t)s->n0));
        FRMC_CUDA(cudaMemcpyAsync(s->d_inv, s->lay.inv.data(), sizeof(int) * (size_t)s->n0, cudaMemcpyHostToDevice, s->stream));
    }
    if (s->isPBC) {
        FRMC_REQUIRE(real && rbasis, FRMC_EINVAL, "a periodic store needs realCoordinates and reciprocalBasisVectors");
        for (int i = 0; i < 9; ++i) s->rbasis[i] = rbasis[i];
        std::vector<float> r4((size_t)s->n0 * 4, 0.f);
        for (int64_t i = 0; i < s->n0; ++i) { r4[4 * i] = real[3 * i]; r4[4 * i + 1] = real[3 * i + 1]; r4[4 * i + 2] = real[3 * i + 2]; }
        if (!s->d_real) FRMC_CUDA(cudaMalloc(&s->d_real, sizeof(float4) * (size_t)s->n0));
        FRMC_CUDA(cudaMemcpyAsync(s->d_real, r4.data(), sizeof(float4) * (size_t)s->n0, cudaMemcpyHostToDevice, s->stream));
    }
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    s->real_valid = true;
    return FRMC_OK;
}

int frmc_store_get_real_coords(frmc_store *s, float *real_out)
{
    FRMC_REQUIRE(s && real_out, FRMC_EINVAL, "NULL argument");
    if (!s->isPBC) return frmc_store_get_coords(s, real_out);       // engine.realCoordinates is engine.boxCoordinates (Engine.py:2253)
    FRMC_REQUIRE(s->real_valid && s->d_real, FRMC_ESTATE, "no valid real coordinates (frmc_store_set_real_coords, then generated runs only)");
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    std::vector<float> r4((size_t)s->n0 * 4);
    FRMC_CUDA(cudaMemcpyAsync(r4.data(), s->d_real, sizeof(float4) * (size_t)s->n0, cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < s->n0; ++i) { real_out[3 * i] = r4[4 * i]; real_out[3 * i + 1] = r4[4 * i + 1]; real_out[3 * i + 2] = r4[4 * i + 2]; }
    return FRMC_OK;
}

int frmc_store_set_groups(frmc_store *s, int n_groups, const int32_t *offsets, const int32_t *indexes)
{
    FRMC_REQUIRE(s && offsets && indexes, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(n_groups >= 1, FRMC_EINVAL, "need at least one group");
    FRMC_REQUIRE(offsets[0] == 0, FRMC_EINVAL, "offsets must start at 0");
    for (int g = 0; g < n_groups; ++g) {
        const int k = offsets[g + 1] - offsets[g];
        FRMC_REQUIRE(k >= 1 && k <= FRMC_MAX_GROUP, FRMC_ELIMIT, "group %d has %d atoms (1..%d)", g, k, FRMC_MAX_GROUP);
        for (int t = offsets[g]; t < offsets[g + 1]; ++t)
            FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < s->n0, FRMC_EINVAL, "group %d: atom index %d outside 0..%lld", g, indexes[t], (long long)s->n0 - 1);
    }
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
    cudaFree(s->d_goff); cudaFree(s->d_gidx); s->d_goff = s->d_gidx = nullptr;
    FRMC_CUDA(cudaMalloc(&s->d_goff, sizeof(int) * (size_t)(n_groups + 1)));
    FRMC_CUDA(cudaMalloc(&s->d_gidx, sizeof(int) * (size_t)offsets[n_groups]));
    FRMC_CUDA(cudaMemcpy(s->d_goff, offsets, sizeof(int) * (size_t)(n_groups + 1), cudaMemcpyHostToDevice));
    FRMC_CUDA(cudaMemcpy(s->d_gidx, indexes, sizeof(int) * (size_t)offsets[n_groups], cudaMemcpyHostToDevice));
    s->n_groups = n_groups;
    return FRMC_OK;
}

int frmc_run_generated(frmc_store *s, int n, uint64_t seed, uint64_t first_counter, float amp_min, float amp_max,
                       const float *variance_sq, float tolerance, float *total_io, float *chi2_out, int32_t *decisions,
                       int32_t *groups_out, float *rand_out, double *device_ms)
{
    FRMC_REQUIRE(s && total_io, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(n >= 1, FRMC_EINVAL, "need at least one proposal");
    FRMC_REQUIRE(amp_min >= 0.f && amp_max > amp_min, FRMC_EINVAL, "bad amplitude range [%g, %g)", (double)amp_min, (double)amp_max);
    FRMC_REQUIRE(s->state == 0, FRMC_ESTATE, "a proposal is staged; accept or reject it first");
    FRMC_REQUIRE(!s->grids.empty() && !s->models.empty(), FRMC_ESTATE, "a run of proposals needs at least one grid and one model");
    for (auto &g : s->grids) FRMC_REQUIRE(g.valid, FRMC_ESTATE, "call frmc_compute_data before proposing moves");
    FRMC_REQUIRE(s->n_groups > 0, FRMC_ESTATE, "no groups (frmc_store_set_groups)");
    FRMC_REQUIRE(s->real_valid, FRMC_ESTATE, "no valid real coordinates (frmc_store_set_real_coords)");
    FRMC_REQUIRE(s->rel2real.empty(), FRMC_ESTATE, "atoms were removed from this store: re-create it before generated runs");
    const int nm = (int)s->models.size();
    FRMC_CUDA(cudaSetDevice(s->dev));
    int rc = flush_pending(s);
    if (rc) return rc;
    if ((rc = sync_models(s))) return rc;
    FRMC_REQUIRE(s->batch_ok && !s->timing, FRMC_ESTATE,
                 "generated runs go through the batch kernel: resident S(Q) slabs, no scale-factor refit schedule, timing off");
    if ((rc = batch_prepare(s))) return rc;
    if (!s->d_bin) {
        FRMC_CUDA(cudaMalloc(&s->d_bin, sizeof(BatchIn)));
        FRMC_CUDA(cudaMalloc(&s->d_gen, sizeof(GenOut)));
    }
    if (s->brand_cap < (size_t)n + 2 * BATCH_MAX_GROUPS) {
        cudaFree(s->d_brand); s->d_brand = nullptr; s->brand_cap = 0;
        FRMC_CUDA(cudaMalloc(&s->d_brand, sizeof(float) * ((size_t)n + 2 * BATCH_MAX_GROUPS)));
        s->brand_cap = (size_t)n + 2 * BATCH_MAX_GROUPS;
    }
    if (s->bout_cap < (size_t)n) {
        cudaFree(s->d_bout_chi2); cudaFree(s->d_bout_dec); s->d_bout_chi2 = nullptr; s->d_bout_dec = nullptr; s->bout_cap = 0;
        FRMC_CUDA(cudaMalloc(&s->d_bout_chi2, sizeof(float) * (size_t)n * FRMC_MAX_MODELS));
        FRMC_CUDA(cudaMalloc(&s->d_bout_dec, sizeof(int) * (size_t)n));
        s->bout_cap = (size_t)n;
    }
    if (s->gout_cap < (size_t)n) {
        cudaFree(s->d_gout); s->d_gout = nullptr; s->gout_cap = 0;
        FRMC_CUDA(cudaMalloc(&s->d_gout, sizeof(int) * (size_t)n));
        s->gout_cap = (size_t)n;
    }
    FRMC_CUDA(cudaMemsetAsync(s->d_brand, 0, sizeof(float) * ((size_t)n + 2 * BATCH_MAX_GROUPS), s->stream));
    BatchDev &bd = s->bdev;
    bd.rand = s->d_brand; bd.out_chi2 = s->d_bout_chi2; bd.out_dec = s->d_bout_dec;
    for (int i = 0; i < FRMC_MAX_MODELS; ++i) bd.var2[i] = (variance_sq && i < nm) ? variance_sq[i] : 1.0f;
    bd.tol = tolerance;
    bd.rand_per_proposal = 1;
    for (int i = 0; i < FRMC_MAX_MODELS; ++i) bd.freq[i] = (i < nm) ? s->models[i].adjust_freq : 0;
    bd.accepted_base = s->accepted;
    BatchRun run0;
    memset(&run0, 0, sizeof(run0));
    run0.total = *total_io;
    for (int m = 0; m < nm; ++m) { run0.cchi2[m] = s->chi2_committed[m]; run0.csf[m] = s->models[m].dev.scale; }
    *s->h_brun = run0;
    FRMC_CUDA(cudaMemcpyAsync(bd.run, s->h_brun, sizeof(BatchRun), cudaMemcpyHostToDevice, s->stream));
    if (device_ms) FRMC_CUDA(cudaEventRecord(s->bev0, s->stream));
    GenParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.seed = seed; gp.first_counter = first_counter; gp.n_total = n; gp.n_groups = s->n_groups;
    gp.goff = s->d_goff; gp.gidx = s->d_gidx; gp.inv = s->d_inv; gp.real = s->d_real;
    for (int i = 0; i < 9; ++i) gp.rb[i] = s->rbasis[i];
    gp.pbc = s->isPBC; gp.amp_min = amp_min; gp.amp_max = amp_max;
    gp.rand_out = s->d_brand; gp.group_out = s->d_gout;
    BatchIn none;
    memset(&none, 0, sizeof(none));
    int done = 0, replans = 0;
    while (done < n) {
        // The window the generated box coordinates must stay in: the geometry mode (the fast minimum image needs every
        // coordinate difference below 1.5) and the rounding margin of the culling boxes are chosen for it.  It is fixed
        // once per layout (and once more if a coordinate ever leaves the fast window); the store's bounds become the
        // window, since an accepted position may be anywhere inside.
        if (s->gen_win_kind == 0 || (s->gen_win_kind == 1 && s->force_general)) {
            float mid[3], half[3];
            for (int c = 0; c < 3; ++c) { mid[c] = 0.5f * (s->lo[c] + s->hi[c]); half[c] = 0.5f * (s->hi[c] - s->lo[c]); }
            const bool fast = s->isPBC && !s->force_general && half[0] <= 0.7449f && half[1] <= 0.7449f && half[2] <= 0.7449f;
            for (int c = 0; c < 3; ++c) {
                const float w = !s->isPBC ? half[c] + std::max(16.0f, 2.0f * half[c]) : (fast ? 0.7449f : std::max(8.0f, half[c] + 4.0f));
                s->gen_win_lo[c] = mid[c] - w; s->gen_win_hi[c] = mid[c] + w;
            }
            s->gen_win_kind = fast ? 1 : 2;
            if (s->isPBC && !fast) s->force_general = true;
        }
        const bool fast = s->gen_win_kind == 1;
        for (int c = 0; c < 3; ++c) {
            gp.win_lo[c] = s->gen_win_lo[c]; gp.win_hi[c] = s->gen_win_hi[c];
            s->lo[c] = std::min(s->lo[c], gp.win_lo[c]); s->hi[c] = std::max(s->hi[c], gp.win_hi[c]);
        }
        const int mode = current_mode(s, nullptr, nullptr);
        const int launches = std::min(64, (n - done + BATCH_MAX_PROPS - 1) / BATCH_MAX_PROPS + 1);
        for (int l = 0; l < launches; ++l) {
            generate_batch_kernel<<<1, 32, 0, s->stream>>>(gp, bd.run, s->d_bin, s->d_gen, s->d_atoms);
            FRMC_LAUNCH_CHECK();
            if ((rc = launch_batch(s, mode, nullptr, 0, true))) return rc;
        }
        FRMC_CUDA(cudaMemcpyAsync(s->h_brun, bd.run, sizeof(BatchRun), cudaMemcpyDeviceToHost, s->stream));
        FRMC_CUDA(cudaStreamSynchronize(s->stream));
        const BatchRun &r = *s->h_brun;
        FRMC_REQUIRE(r.n_done >= done && r.n_done <= n, FRMC_ECUDA, "generated run went backwards (done %d -> %d of %d)", done, r.n_done, n);
        if (r.gen_state == 3) {
            // a generated coordinate left the window: general minimum image from now on (exact for any coordinate)
            FRMC_REQUIRE(fast && ++replans <= 2, FRMC_ELIMIT,
                         "generated coordinates drifted out of the store's coordinate window; re-wrap them (frmc_store_set_coords)");
            s->force_general = true;
            s->h_brun->gen_state = 0;
            FRMC_CUDA(cudaMemcpyAsync(bd.run, s->h_brun, sizeof(BatchRun), cudaMemcpyHostToDevice, s->stream));
        } else {
            FRMC_REQUIRE(r.n_done > done, FRMC_ECUDA, "generated run made no progress (done %d of %d)", done, n);
        }
        done = r.n_done;
    }
    if (device_ms) {
        FRMC_CUDA(cudaEventRecord(s->bev1, s->stream));
        FRMC_CUDA(cudaEventSynchronize(s->bev1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->bev0, s->bev1);
        *device_ms = ms;
    }
    if (chi2_out) FRMC_CUDA(cudaMemcpyAsync(chi2_out, s->d_bout_chi2, sizeof(float) * (size_t)n * nm, cudaMemcpyDeviceToHost, s->stream));
    if (decisions) FRMC_CUDA(cudaMemcpyAsync(decisions, s->d_bout_dec, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    if (groups_out) FRMC_CUDA(cudaMemcpyAsync(groups_out, s->d_gout, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    if (rand_out) FRMC_CUDA(cudaMemcpyAsync(rand_out, s->d_brand, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    const BatchRun &r = *s->h_brun;
    *total_io = r.total;
    if (r.n_accepted > 0)
        for (int m = 0; m < nm; ++m) {
            s->chi2_committed[m] = r.cchi2[m]; s->chi2_staged[m] = r.cchi2[m];
            if (s->models[m].adjust_freq > 0) { s->models[m].dev.scale = r.csf[m]; s->models[m].sf_staged = r.csf[m]; }
        }
    s->accepted += (unsigned long long)r.n_accepted;
    s->batch_rounds += (unsigned long long)r.rounds;
    s->batch_proposals += (unsigned long long)n;
    return FRMC_OK;
}

int frmc_store_batch_stats(frmc_store *s, uint64_t *launches, uint64_t *rounds, uint64_t *proposals)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    if (launches) *launches = s->batch_launches;
    if (rounds) *rounds = s->batch_rounds;
    if (proposals) *proposals = s->batch_proposals;
    return FRMC_OK;
}

int frmc_store_batch_stamps(frmc_store *s, int64_t *out, int n)
{
    FRMC_REQUIRE(s && out && n >= 1 && n <= BATCH_STAMP_TOTAL, FRMC_EINVAL, "bad arguments");
    FRMC_REQUIRE(s->d_bstamps, FRMC_ESTATE, "no batch timeline recorded (set FRMC_BATCH_STAMPS=1 before the first run)");
    FRMC_CUDA(cudaSetDevice(s->dev));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    FRMC_CUDA(cudaMemcpy(out, s->d_bstamps, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    return FRMC_OK;
}

int frmc_store_committed_chi2(frmc_store *s, float *chi2)
{
    FRMC_REQUIRE(s && chi2, FRMC_EINVAL, "NULL argument");
    for (size_t i = 0; i < s->models.size(); ++i) chi2[i] = s->chi2_committed[i];
    return FRMC_OK;
}

int frmc_store_replay_proposal(frmc_store *s, int reps, double *ms_per_launch)
{
    FRMC_REQUIRE(s && reps >= 1 && ms_per_launch, FRMC_EINVAL, "bad arguments");
    FRMC_REQUIRE(s->state == 1, FRMC_ESTATE, "stage a proposal with frmc_propose first");
    FRMC_CUDA(cudaSetDevice(s->dev));
    const int mode = current_mode(s, s->prop_lo, s->prop_hi);
    const bool timing = s->timing;
    s->timing = false;
    // back-to-back launches are the point of the replay: it always runs the launch-per-proposal kernels
    const bool persist = s->persist_enabled;
    { int prc = stop_persistent(s); if (prc) return prc; }
    s->persist_enabled = false;
    cudaEvent_t e0, e1;
    FRMC_CUDA(cudaEventCreate(&e0));
    FRMC_CUDA(cudaEventCreate(&e1));
    int rc = FRMC_OK;
    FRMC_CUDA(cudaEventRecord(e0, s->stream));
    for (int i = 0; i < reps && rc == FRMC_OK; ++i) { ++s->seq_expected; rc = launch_propose(s, mode); }
    FRMC_CUDA(cudaEventRecord(e1, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_launch = (double)ms / reps;
    // the replays stacked reps extra copies of the delta: clear and stage the proposal once more
    GridSet gs = make_gridset(s);
    clear_delta_kernel<<<launch_cells_grid(s), 256, 0, s->stream>>>(gs);
    FRMC_LAUNCH_CHECK();
    ++s->seq_expected;
    if (rc == FRMC_OK) rc = launch_propose(s, mode);
    if (rc == FRMC_OK) rc = wait_epilogue(s);
    s->timing = timing;
    s->persist_enabled = persist;
    return rc;
}

int frmc_export_data(frmc_store *s, int grid, float *hintra, float *hinter)
{
    FRMC_REQUIRE(s && hintra && hinter, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(grid >= 0 && grid < (int)s->grids.size(), FRMC_EINVAL, "unknown grid %d", grid);
    FRMC_CUDA(cudaSetDevice(s->dev));
    { int frc = flush_pending(s); if (frc) return frc; }
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
    { int frc = flush_pending(s); if (frc) return frc; }
    ModelHost &m = s->models[model];
    FRMC_CUDA(cudaMemcpyAsync(out, staged ? m.dev.total : m.total_committed, sizeof(float) * m.dev.n_out,
                              cudaMemcpyDeviceToHost, s->stream));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    return FRMC_OK;
}

int frmc_store_set_timing(frmc_store *s, int on)
{
    FRMC_REQUIRE(s, FRMC_EINVAL, "NULL store");
    { int frc = flush_pending(s); if (frc) return frc; }
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

int frmc_store_debug_stamps(frmc_store *s, int64_t *out, int n)
{
    FRMC_REQUIRE(s && out && n >= 1 && n <= 16 * FRMC_MAX_MODELS, FRMC_EINVAL, "bad arguments");
    FRMC_CUDA(cudaSetDevice(s->dev));
    FRMC_CUDA(cudaStreamSynchronize(s->stream));
    FRMC_CUDA(cudaMemcpy(out, s->d_stamps, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    return FRMC_OK;
}

uint64_t frmc_store_edge_overflow(frmc_store *s)
{
    if (!s) return 0;
    stop_persistent(s);
    unsigned long long ov = 0;
    cudaSetDevice(s->dev);
    cudaMemcpyAsync(&ov, s->d_overflow, sizeof(ov), cudaMemcpyDeviceToHost, s->stream);
    cudaStreamSynchronize(s->stream);
    return ov;
}

uint64_t frmc_store_swept_pairs(frmc_store *s)
{
    if (!s) return 0;
    stop_persistent(s);
    unsigned long long blocks = 0;
    cudaSetDevice(s->dev);
    cudaMemcpyAsync(&blocks, s->d_overflow + 1, sizeof(blocks), cudaMemcpyDeviceToHost, s->stream);
    cudaStreamSynchronize(s->stream);
    return blocks * 256ull;               // (32 x 8)-record chunks
}

}  // extern "C"
