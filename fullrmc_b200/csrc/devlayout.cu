// devlayout.cu -- the element-sorted, k-d ordered atom store (layout.h) built ON THE DEVICE from the caller's raw
// arrays (boxCoords [N,3], elementIndex [N], moleculeIndex [N]).
//
// The stateless entry points of fullhist.cu receive host arrays on every call (an Engine calling compute_data; the
// e2e leg of bench.py).  Ordering them on the host (build_layout: nth_element over all atoms, 18 ms for 10^6 atoms on
// 16 cores) cost more than the histogram itself; here the host only walks the tree SHAPE, which depends on nothing
// but the element counts, and the device does the data-dependent work:
//
//   dl_count_kernel     per 1024-atom chunk: atoms per element, coordinate bounds, finiteness, element range check
//   dl_scan_kernel      exclusive scan of the chunk counts per element (stable positions), totals   -> host (1 sync)
//   dl_gather_kernel    stable counting sort by element into 16-byte points {reduced x, y, z, original index}
//   dl_split_kernel     one CTA per tree node, level by level: the node's points are split at record k (a multiple of
//                       1024 / 256 / 32, layout.h) along the longest axis of the node's box -- a 2048-bin histogram of
//                       the coordinate over the box extent finds the bin holding the k-th point, a stable partition
//                       (ballot prefix sums, chunks in order) moves lower bins left and higher bins right, the
//                       boundary bin is cut by arrival order.  Points inside one bin (1/2048 of the node's extent) may
//                       land on either side: the boxes of the two halves then overlap by that much, which the exact
//                       culling test of fullhist.cu (bounding boxes of the records themselves) never notices.
//   dl_leaf_kernel      nodes of <= 1024 points: the remaining levels in shared memory (exact rank sort on
//                       (coordinate, original index) per level), then the final records: {raw x, y, z, meta},
//                       original index, padding with NaN records
//
// Every step is deterministic (no arrival-order atomics decide a position), so two processes given the same arrays
// build the same order -- the multi-rank shards of one histogram rely on that.  Any order yields the same histogram
// (integer counts; the ordered [a,b] slot comes from the original indexes); the order only decides how many block
// pairs the sweep can skip.
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

struct DlPoint { float x, y, z; uint32_t idx; };

struct DlStats {                 // device -> host after the counting pass
    int count[FRMC_MAX_ELEMENTS];
    int lo[3], hi[3];            // order-preserving integer images of the coordinate bounds
    int not_finite, bad_element;
    int mol_not_direct;          // a molecule index outside [0, 2^24 - 1): the host ranks them
    int mol_decreasing;          // moleculeIndex is not non-decreasing: molecules may be scattered over the index range
    int mol_repeats;             // two consecutive atoms share a molecule: molecules of more than one atom exist
};

struct DlTask {                  // one tree node (host-built: the shape depends only on the element counts)
    int start, len;              // range in the compact element-major point array
    int k;                       // split: records [start, start+k) go left; leaf: unused
    int id, left, right;         // box slots of the node and of its children
    int src;                     // which of the two point buffers holds the node's points
    int elem;                    // element (leaf: padded position = position + shift[elem])
};

static const int DL_CHUNK = 1024;        // atoms per CTA of the counting / gather passes
static const int DL_BINS = 2048;
static const int DL_LEAF = 1024;         // nodes up to this many points finish in shared memory

__device__ __forceinline__ int dl_float_to_ordered(float f)
{
    const int i = __float_as_int(f);
    return (i >= 0) ? i : (i ^ 0x7FFFFFFF);
}
static float dl_ordered_to_float(int i)
{
    const int b = (i >= 0) ? i : (i ^ 0x7FFFFFFF);
    float f;
    memcpy(&f, &b, 4);
    return f;
}

__global__ void __launch_bounds__(256) dl_count_kernel(const float *__restrict__ coords, const int32_t *__restrict__ el,
                                                       const int32_t *__restrict__ mol, long long n, int nEl,
                                                       int *__restrict__ chunk_cnt, DlStats *__restrict__ stats)
{
    __shared__ int s_cnt[FRMC_MAX_ELEMENTS];
    __shared__ int s_lo[3], s_hi[3], s_flags[2];
    const int tid = threadIdx.x;
    if (tid < FRMC_MAX_ELEMENTS) s_cnt[tid] = 0;
    if (tid < 3) { s_lo[tid] = 0x7FFFFFFF; s_hi[tid] = (int)0x80000000; }
    if (tid < 2) s_flags[tid] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * DL_CHUNK;
    int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    int nf = 0, bad = 0, mflags = 0;
    for (int r = 0; r < DL_CHUNK / 256; ++r) {
        const long long i = base + r * 256 + tid;
        if (i < n) {
            const int e = el[i];
            if (e < 0 || e >= nEl) bad = 1; else atomicAdd(&s_cnt[e], 1);
            const int m = mol[i];
            if (m < 0 || m >= 0x00FFFFFF) mflags |= 1;
            if (i > 0) {
                const int mp = mol[i - 1];
                if (m < mp) mflags |= 2;
                if (m == mp) mflags |= 4;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = coords[3 * i + c];
                if (!isfinite(v)) { nf = 1; continue; }
                const int o = dl_float_to_ordered(v);
                lo[c] = min(lo[c], o); hi[c] = max(hi[c], o);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        lo[c] = __reduce_min_sync(0xffffffffu, lo[c]);
        hi[c] = __reduce_max_sync(0xffffffffu, hi[c]);
    }
    nf = __any_sync(0xffffffffu, nf); bad = __any_sync(0xffffffffu, bad);
    mflags = __reduce_or_sync(0xffffffffu, mflags);
    if ((tid & 31) == 0 && mflags) {
        if (mflags & 1) stats->mol_not_direct = 1;
        if (mflags & 2) stats->mol_decreasing = 1;
        if (mflags & 4) stats->mol_repeats = 1;
    }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { atomicMin(&s_lo[c], lo[c]); atomicMax(&s_hi[c], hi[c]); }
        if (nf) s_flags[0] = 1;
        if (bad) s_flags[1] = 1;
    }
    __syncthreads();
    if (tid < nEl) chunk_cnt[(size_t)blockIdx.x * FRMC_MAX_ELEMENTS + tid] = s_cnt[tid];
    if (tid < 3) { atomicMin(&stats->lo[tid], s_lo[tid]); atomicMax(&stats->hi[tid], s_hi[tid]); }
    if (tid == 0) { if (s_flags[0]) stats->not_finite = 1; if (s_flags[1]) stats->bad_element = 1; }
}

// exclusive scan over the chunks, per element: chunk_cnt -> chunk_off (in place), totals -> stats->count
__global__ void __launch_bounds__(1024) dl_scan_kernel(int *__restrict__ chunk_cnt, int n_chunks, int nEl, DlStats *__restrict__ stats)
{
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int per = (n_chunks + 1023) / 1024;
    const int c0 = min(n_chunks, t * per), c1 = min(n_chunks, c0 + per);
    for (int e = 0; e < nEl; ++e) {
        int sum = 0;
        for (int c = c0; c < c1; ++c) sum += chunk_cnt[(size_t)c * FRMC_MAX_ELEMENTS + e];
        part[t] = sum;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int v = (t >= o) ? part[t - o] : 0;
            __syncthreads();
            part[t] += v;
            __syncthreads();
        }
        int run = part[t] - sum;
        for (int c = c0; c < c1; ++c) {
            const int v = chunk_cnt[(size_t)c * FRMC_MAX_ELEMENTS + e];
            chunk_cnt[(size_t)c * FRMC_MAX_ELEMENTS + e] = run;
            run += v;
        }
        if (t == 1023) stats->count[e] = part[t];
        __syncthreads();
    }
}

struct DlElemBase { int base[FRMC_MAX_ELEMENTS]; };     // compact start of every element's points

// stable counting sort by element: point k of element e = the k-th atom of that element in original order
__global__ void __launch_bounds__(256) dl_gather_kernel(const float *__restrict__ coords, const int32_t *__restrict__ el, long long n, int nEl,
                                                        int pbc, const int *__restrict__ chunk_off, DlElemBase eb, DlPoint *__restrict__ pts)
{
    __shared__ int s_run[FRMC_MAX_ELEMENTS];            // atoms of each element already placed by this CTA
    __shared__ int s_warp[8][FRMC_MAX_ELEMENTS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < FRMC_MAX_ELEMENTS) s_run[tid] = (tid < nEl) ? chunk_off[(size_t)blockIdx.x * FRMC_MAX_ELEMENTS + tid] + eb.base[tid] : 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * DL_CHUNK;
    for (int r = 0; r < DL_CHUNK / 256; ++r) {
        const long long i = base + r * 256 + tid;
        const int e = (i < n) ? el[i] : -1;
        int rank_in_warp = 0;
        for (int q = 0; q < nEl; ++q) {
            const unsigned m = __ballot_sync(0xffffffffu, e == q);
            if (e == q) rank_in_warp = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) s_warp[warp][q] = __popc(m);
        }
        __syncthreads();
        if (e >= 0) {
            int before = s_run[e];
            for (int w = 0; w < warp; ++w) before += s_warp[w][e];
            DlPoint p;
            const float v[3] = {coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]};
            float f[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) f[c] = isfinite(v[c]) ? (pbc ? (v[c] - floorf(v[c])) : v[c]) : 0.f;
            p.x = f[0]; p.y = f[1]; p.z = f[2]; p.idx = (uint32_t)i;
            pts[before + rank_in_warp] = p;
        }
        __syncthreads();
        if (tid < nEl) {
            int add = 0;
            for (int w = 0; w < 8; ++w) add += s_warp[w][tid];
            s_run[tid] += add;
        }
        __syncthreads();
    }
}

struct DlBox { float lo[3], hi[3]; };

// one CTA per node: split [start, start+len) at record k along the longest axis of the node's box
__global__ void __launch_bounds__(1024) dl_split_kernel(const DlTask *__restrict__ tasks, DlPoint *__restrict__ buf0, DlPoint *__restrict__ buf1,
                                                        DlBox *__restrict__ boxes)
{
    __shared__ int hist[DL_BINS];
    __shared__ int s_scan[1024];
    __shared__ int s_wl[32], s_wm[32], s_wr[32];
    __shared__ int s_qcut, s_less, s_eq;
    const DlTask T = tasks[blockIdx.x];
    const DlPoint *__restrict__ src = (T.src ? buf1 : buf0) + T.start;
    DlPoint *__restrict__ dst = (T.src ? buf0 : buf1) + T.start;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DlBox B = boxes[T.id];
    int ax = 0;
    if (B.hi[1] - B.lo[1] > B.hi[ax] - B.lo[ax]) ax = 1;
    if (B.hi[2] - B.lo[2] > B.hi[ax] - B.lo[ax]) ax = 2;
    const float lo = B.lo[ax], ext = B.hi[ax] - B.lo[ax];
    const float scale = (ext > 0.f) ? (float)DL_BINS / ext : 0.f;
    auto bin_of = [&](const DlPoint &p) {
        const float v = (ax == 0) ? p.x : (ax == 1) ? p.y : p.z;
        return max(0, min(DL_BINS - 1, (int)((v - lo) * scale)));
    };
    for (int b = tid; b < DL_BINS; b += 1024) hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < T.len; i += 1024) atomicAdd(&hist[bin_of(src[i])], 1);
    __syncthreads();
    // the bin holding record k (1-based rank k): less = records in lower bins < k <= less + eq
    {
        const int a = hist[2 * tid], b = hist[2 * tid + 1];
        s_scan[tid] = a + b;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int v = (tid >= o) ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int before = s_scan[tid] - (a + b);
        if (before < T.k && T.k <= before + a) { s_qcut = 2 * tid; s_less = before; s_eq = a; }
        else if (before + a < T.k && T.k <= before + a + b) { s_qcut = 2 * tid + 1; s_less = before + a; s_eq = b; }
        __syncthreads();
    }
    const int qcut = s_qcut, less = s_less, eq = s_eq;
    const int r = T.k - less;                       // records of the boundary bin that go left (arrival order)
    int doneL = 0, doneM = 0, doneR = 0;            // placed so far (CTA-uniform)
    for (int base = 0; base < T.len; base += 1024) {
        const int i = base + tid;
        DlPoint p;
        int cls = 3;                                // 0 lower bins, 1 boundary bin, 2 higher bins
        if (i < T.len) {
            p = src[i];
            const int q = bin_of(p);
            cls = (q < qcut) ? 0 : (q == qcut) ? 1 : 2;
        }
        const unsigned mL = __ballot_sync(0xffffffffu, cls == 0), mM = __ballot_sync(0xffffffffu, cls == 1),
                       mR = __ballot_sync(0xffffffffu, cls == 2);
        if (lane == 0) { s_wl[warp] = __popc(mL); s_wm[warp] = __popc(mM); s_wr[warp] = __popc(mR); }
        __syncthreads();
        int pl = 0, pm = 0, pr = 0, tl = 0, tm = 0, tr = 0;
        for (int w = 0; w < 32; ++w) {
            const int a = s_wl[w], b = s_wm[w], c = s_wr[w];
            if (w < warp) { pl += a; pm += b; pr += c; }
            tl += a; tm += b; tr += c;
        }
        const unsigned lt = (1u << lane) - 1u;
        if (cls == 0) {
            dst[doneL + pl + __popc(mL & lt)] = p;
        } else if (cls == 1) {
            const int m = doneM + pm + __popc(mM & lt);
            dst[(m < r) ? (less + m) : (T.k + (m - r))] = p;
        } else if (cls == 2) {
            dst[T.k + (eq - r) + doneR + pr + __popc(mR & lt)] = p;
        }
        doneL += tl; doneM += tm; doneR += tr;
        __syncthreads();
    }
    if (tid == 0) {
        const float cut = lo + ((float)qcut + 0.5f) * ((scale > 0.f) ? 1.0f / scale : 0.f);
        DlBox L = B, R = B;
        L.hi[ax] = cut; R.lo[ax] = cut;
        boxes[T.left] = L; boxes[T.right] = R;
    }
}

// ---- the same split for LARGE nodes, spread over many CTAs ---------------------------------------------------------
// One CTA per node leaves the top of the tree (5 nodes of 200 000 points at cfg5, then 10, 20, 40) to a handful of SMs:
// 1.3 of the 1.5 ms the splits of 10^6 atoms took.  Nodes of at least DL_WIDE_MIN points are cut into chunks of DL_WCHUNK
// points, one CTA each, in three launches per level:
//   dl_whist_kernel     the chunk's 2048-bin histogram -> its private copy in global memory + integer atomics into the
//                       node's histogram
//   dl_wcut_kernel      one CTA per node: the bin holding record k (as dl_split_kernel), then from the chunks' private
//                       histograms the records each chunk sends left / into the boundary bin / right, scanned over the
//                       chunks IN ORDER -- the placement stays the stable partition of the single-CTA kernel, record for
//                       record -- and the children's boxes
//   dl_wscatter_kernel  the chunk's records to their places (ballot prefix sums inside the chunk, chunk offsets from the scan)
static const int DL_WCHUNK = 4096;
static const int DL_WIDE_MIN = 20000;

struct DlWChunk { int task, first, len; };                 // task of the level's wide list, range inside the node
struct DlWCut { int qcut, less, eq, r, ax; float lo, scale; };

__device__ __forceinline__ int dl_bin_of(const DlPoint &p, int ax, float lo, float scale)
{
    const float v = (ax == 0) ? p.x : (ax == 1) ? p.y : p.z;
    return max(0, min(DL_BINS - 1, (int)((v - lo) * scale)));
}

__device__ __forceinline__ void dl_axis_of(const DlBox &B, int &ax, float &lo, float &scale)
{
    ax = 0;
    if (B.hi[1] - B.lo[1] > B.hi[ax] - B.lo[ax]) ax = 1;
    if (B.hi[2] - B.lo[2] > B.hi[ax] - B.lo[ax]) ax = 2;
    lo = B.lo[ax];
    const float ext = B.hi[ax] - B.lo[ax];
    scale = (ext > 0.f) ? (float)DL_BINS / ext : 0.f;
}

__global__ void __launch_bounds__(1024) dl_whist_kernel(const DlTask *__restrict__ tasks, const DlWChunk *__restrict__ chunks,
                                                        const DlPoint *__restrict__ buf0, const DlPoint *__restrict__ buf1,
                                                        const DlBox *__restrict__ boxes, int *__restrict__ chunk_hist, int *__restrict__ node_hist)
{
    __shared__ int hist[DL_BINS];
    const DlWChunk C = chunks[blockIdx.x];
    const DlTask T = tasks[C.task];
    const DlPoint *__restrict__ src = (T.src ? buf1 : buf0) + T.start + C.first;
    int ax; float lo, scale;
    dl_axis_of(boxes[T.id], ax, lo, scale);
    const int tid = threadIdx.x;
    for (int b = tid; b < DL_BINS; b += 1024) hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < C.len; i += 1024) atomicAdd(&hist[dl_bin_of(src[i], ax, lo, scale)], 1);
    __syncthreads();
    for (int b = tid; b < DL_BINS; b += 1024) {
        const int h = hist[b];
        chunk_hist[(size_t)blockIdx.x * DL_BINS + b] = h;
        if (h) atomicAdd(&node_hist[(size_t)C.task * DL_BINS + b], h);
    }
}

// chunk_first[t] .. chunk_first[t+1]: the chunks of wide task t, in node order; chunk_off[c] = {left, boundary, right}
// records of the node's earlier chunks
__global__ void __launch_bounds__(1024) dl_wcut_kernel(const DlTask *__restrict__ tasks, const int *__restrict__ chunk_first,
                                                       const int *__restrict__ chunk_hist, int *__restrict__ node_hist,
                                                       DlBox *__restrict__ boxes, DlWCut *__restrict__ cuts, int3 *__restrict__ chunk_off,
                                                       const DlWChunk *__restrict__ chunks)
{
    __shared__ int s_scan[1024];
    __shared__ int s_qcut, s_less, s_eq;
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DlTask T = tasks[t];
    const DlBox B = boxes[T.id];
    int ax; float lo, scale;
    dl_axis_of(B, ax, lo, scale);
    int *hist = node_hist + (size_t)t * DL_BINS;
    {
        const int a = hist[2 * tid], b = hist[2 * tid + 1];
        hist[2 * tid] = 0; hist[2 * tid + 1] = 0;                 // re-armed for the next level
        s_scan[tid] = a + b;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int v = (tid >= o) ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int before = s_scan[tid] - (a + b);
        if (before < T.k && T.k <= before + a) { s_qcut = 2 * tid; s_less = before; s_eq = a; }
        else if (before + a < T.k && T.k <= before + a + b) { s_qcut = 2 * tid + 1; s_less = before + a; s_eq = b; }
        __syncthreads();
    }
    const int qcut = s_qcut, less = s_less, eq = s_eq;
    const int c0 = chunk_first[t], c1 = chunk_first[t + 1];
    // per chunk: records below the boundary bin and inside it
    for (int c = c0 + warp; c < c1; c += 32) {
        const int *h = chunk_hist + (size_t)c * DL_BINS;
        int below = 0;
        for (int b = lane; b < qcut; b += 32) below += h[b];
        below = __reduce_add_sync(0xffffffffu, below);
        if (lane == 0) chunk_off[c] = make_int3(below, h[qcut], chunks[c].len - below - h[qcut]);
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan over the node's chunks in order, 32 at a time
        int runL = 0, runM = 0, runR = 0;
        for (int base = c0; base < c1; base += 32) {
            const int c = base + lane;
            int3 v = (c < c1) ? chunk_off[c] : make_int3(0, 0, 0);
            int iL = v.x, iM = v.y, iR = v.z;
            for (int o = 1; o < 32; o <<= 1) {
                const int a = __shfl_up_sync(0xffffffffu, iL, o), b = __shfl_up_sync(0xffffffffu, iM, o), d = __shfl_up_sync(0xffffffffu, iR, o);
                if (lane >= o) { iL += a; iM += b; iR += d; }
            }
            if (c < c1) chunk_off[c] = make_int3(runL + iL - v.x, runM + iM - v.y, runR + iR - v.z);
            runL += __shfl_sync(0xffffffffu, iL, 31); runM += __shfl_sync(0xffffffffu, iM, 31); runR += __shfl_sync(0xffffffffu, iR, 31);
        }
    }
    if (tid == 0) {
        DlWCut cut;
        cut.qcut = qcut; cut.less = less; cut.eq = eq; cut.r = T.k - less; cut.ax = ax; cut.lo = lo; cut.scale = scale;
        cuts[t] = cut;
        const float at = lo + ((float)qcut + 0.5f) * ((scale > 0.f) ? 1.0f / scale : 0.f);
        DlBox L = B, R = B;
        L.hi[ax] = at; R.lo[ax] = at;
        boxes[T.left] = L; boxes[T.right] = R;
    }
}

__global__ void __launch_bounds__(1024) dl_wscatter_kernel(const DlTask *__restrict__ tasks, const DlWChunk *__restrict__ chunks,
                                                           const DlWCut *__restrict__ cuts, const int3 *__restrict__ chunk_off,
                                                           DlPoint *__restrict__ buf0, DlPoint *__restrict__ buf1)
{
    __shared__ int s_wl[32], s_wm[32], s_wr[32];
    const DlWChunk C = chunks[blockIdx.x];
    const DlTask T = tasks[C.task];
    const DlWCut cut = cuts[C.task];
    const int3 off = chunk_off[blockIdx.x];
    const DlPoint *__restrict__ src = (T.src ? buf1 : buf0) + T.start + C.first;
    DlPoint *__restrict__ dst = (T.src ? buf0 : buf1) + T.start;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int doneL = off.x, doneM = off.y, doneR = off.z;
    for (int base = 0; base < C.len; base += 1024) {
        const int i = base + tid;
        DlPoint p;
        int cls = 3;
        if (i < C.len) {
            p = src[i];
            const int q = dl_bin_of(p, cut.ax, cut.lo, cut.scale);
            cls = (q < cut.qcut) ? 0 : (q == cut.qcut) ? 1 : 2;
        }
        const unsigned mL = __ballot_sync(0xffffffffu, cls == 0), mM = __ballot_sync(0xffffffffu, cls == 1),
                       mR = __ballot_sync(0xffffffffu, cls == 2);
        if (lane == 0) { s_wl[warp] = __popc(mL); s_wm[warp] = __popc(mM); s_wr[warp] = __popc(mR); }
        __syncthreads();
        int pl = 0, pm = 0, pr = 0, tl = 0, tm = 0, tr = 0;
        for (int w = 0; w < 32; ++w) {
            const int a = s_wl[w], b = s_wm[w], c = s_wr[w];
            if (w < warp) { pl += a; pm += b; pr += c; }
            tl += a; tm += b; tr += c;
        }
        const unsigned lt = (1u << lane) - 1u;
        if (cls == 0) {
            dst[doneL + pl + __popc(mL & lt)] = p;
        } else if (cls == 1) {
            const int m = doneM + pm + __popc(mM & lt);
            dst[(m < cut.r) ? (cut.less + m) : (T.k + (m - cut.r))] = p;
        } else if (cls == 2) {
            dst[T.k + (cut.eq - cut.r) + doneR + pr + __popc(mR & lt)] = p;
        }
        doneL += tl; doneM += tm; doneR += tr;
        __syncthreads();
    }
}

struct DlShift { int shift[FRMC_MAX_ELEMENTS]; };      // padded position - compact position, per element

// one CTA per node of <= DL_LEAF points: the remaining k-d levels in shared memory (a node of len points is split at
// k = (ceil(len / unit) / 2) * unit, unit = 256 above 256 points, 32 above 32; kd_order of fullhist.cu has the same
// rule), then the final records.  A level sorts every sub-node on (coordinate along its longest box axis, original
// index) by rank counting -- exact and order-independent.
__global__ void __launch_bounds__(DL_LEAF) dl_leaf_kernel(const DlTask *__restrict__ tasks, const DlPoint *__restrict__ buf0,
                                                          const DlPoint *__restrict__ buf1, const DlBox *__restrict__ boxes,
                                                          const float *__restrict__ coords, const int32_t *__restrict__ molkey,
                                                          DlShift sh, float4 *__restrict__ atoms, uint32_t *__restrict__ orig, int fast)
{
    __shared__ DlPoint pa[DL_LEAF], pb[DL_LEAF];
    __shared__ int n_start[64], n_len[64], n_next_start[64], n_next_len[64];
    __shared__ DlBox n_box[64], n_next_box[64];
    __shared__ int s_nodes, s_next_nodes, s_any;
    const DlTask T = tasks[blockIdx.x];
    const DlPoint *__restrict__ src = (T.src ? buf1 : buf0) + T.start;
    const int tid = threadIdx.x;
    if (tid < T.len) pa[tid] = src[tid];
    if (tid == 0) { s_nodes = 1; n_start[0] = 0; n_len[0] = T.len; n_box[0] = boxes[T.id]; }
    __syncthreads();
    DlPoint *cur = pa, *nxt = pb;
    if (T.len == DL_LEAF && fast) {
        // A full leaf (all but the last leaf of an element): its sub-nodes are aligned runs of 512 / 256 / 128 / 64
        // points, so every level is a SEGMENTED BITONIC SORT of 64-bit words (order-preserving image of the coordinate
        // along the node's longest axis, original index) with the point's slot as payload -- 185 compare-exchange
        // steps for the five levels against 1984 rank-counting iterations per thread, and the same total order, hence
        // the same records in the same places.  The points stay where they are; only (word, slot) pairs move.
        unsigned long long *word = reinterpret_cast<unsigned long long *>(pb);
        int *slot = reinterpret_cast<int *>(pb) + 2 * DL_LEAF;
        slot[tid] = tid;
        __syncthreads();
        for (int level = 0; level < 5; ++level) {
            const int S = DL_LEAF >> level, v = tid / S;
            {
                const DlBox B = n_box[v];
                int ax = 0;
                if (B.hi[1] - B.lo[1] > B.hi[ax] - B.lo[ax]) ax = 1;
                if (B.hi[2] - B.lo[2] > B.hi[ax] - B.lo[ax]) ax = 2;
                const DlPoint p = pa[slot[tid]];
                const float c = __fadd_rn((ax == 0) ? p.x : (ax == 1) ? p.y : p.z, 0.0f);      // -0 and +0 compare equal: one image
                uint32_t u = __float_as_uint(c);
                u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
                word[tid] = ((unsigned long long)u << 32) | p.idx;
            }
            __syncthreads();
            for (int size = 2; size <= S; size <<= 1)
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    const int j = tid ^ stride;
                    if (j > tid) {
                        const bool up = (size == S) || ((tid & size) == 0);
                        const unsigned long long wi = word[tid], wj = word[j];
                        if ((wi > wj) == up) {
                            word[tid] = wj; word[j] = wi;
                            const int si = slot[tid]; slot[tid] = slot[j]; slot[j] = si;
                        }
                    }
                    __syncthreads();
                }
            if (tid < (1 << level)) {
                const DlBox B = n_box[tid];
                int ax = 0;
                if (B.hi[1] - B.lo[1] > B.hi[ax] - B.lo[ax]) ax = 1;
                if (B.hi[2] - B.lo[2] > B.hi[ax] - B.lo[ax]) ax = 2;
                const DlPoint c = pa[slot[tid * S + S / 2]];
                const float cut = (ax == 0) ? c.x : (ax == 1) ? c.y : c.z;
                DlBox L = B, R = B;
                L.hi[ax] = cut; R.lo[ax] = cut;
                n_next_box[2 * tid] = L; n_next_box[2 * tid + 1] = R;
            }
            __syncthreads();
            if (tid < (2 << level)) n_box[tid] = n_next_box[tid];
            __syncthreads();
        }
        const DlPoint p = pa[slot[tid]];
        const long long i = p.idx;
        const int pos = T.start + tid + sh.shift[T.elem];
        const uint32_t meta = ((uint32_t)molkey[i] << 8) | (uint32_t)T.elem;
        atoms[pos] = make_float4(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], __uint_as_float(meta));
        orig[pos] = (uint32_t)i;
        return;
    }
    for (int level = 0; level < 8; ++level) {
        // my node: the one whose range holds position tid
        const int nn = s_nodes;
        int me = -1;
        for (int v = 0; v < nn; ++v)
            if (tid >= n_start[v] && tid < n_start[v] + n_len[v]) me = v;
        if (tid == 0) { s_next_nodes = 0; s_any = 0; }
        __syncthreads();
        if (tid < T.len) {
            const int s0 = n_start[me], ln = n_len[me];
            if (ln > 32) {
                const DlBox B = n_box[me];
                int ax = 0;
                if (B.hi[1] - B.lo[1] > B.hi[ax] - B.lo[ax]) ax = 1;
                if (B.hi[2] - B.lo[2] > B.hi[ax] - B.lo[ax]) ax = 2;
                const DlPoint p = cur[tid];
                const float key = (ax == 0) ? p.x : (ax == 1) ? p.y : p.z;
                int rank = 0;
                for (int j = 0; j < ln; ++j) {
                    const DlPoint o = cur[s0 + j];
                    const float ok = (ax == 0) ? o.x : (ax == 1) ? o.y : o.z;
                    rank += (ok < key || (ok == key && o.idx < p.idx)) ? 1 : 0;
                }
                nxt[s0 + rank] = p;
                s_any = 1;
            } else {
                nxt[tid] = cur[tid];
            }
        }
        __syncthreads();
        if (!s_any) break;                               // every node is down to <= 32 points
        if (tid == 0) {
            int out = 0;
            for (int v = 0; v < nn; ++v) {
                const int s0 = n_start[v], ln = n_len[v];
                if (ln <= 32) { n_next_start[out] = s0; n_next_len[out] = ln; n_next_box[out] = n_box[v]; ++out; continue; }
                const int unit = (ln > 256) ? 256 : 32;
                const int k = ((ln + unit - 1) / unit / 2) * unit;
                const DlBox B = n_box[v];
                int ax = 0;
                if (B.hi[1] - B.lo[1] > B.hi[ax] - B.lo[ax]) ax = 1;
                if (B.hi[2] - B.lo[2] > B.hi[ax] - B.lo[ax]) ax = 2;
                const DlPoint c = nxt[s0 + k];
                const float cut = (ax == 0) ? c.x : (ax == 1) ? c.y : c.z;
                DlBox L = B, R = B;
                L.hi[ax] = cut; R.lo[ax] = cut;
                n_next_start[out] = s0; n_next_len[out] = k; n_next_box[out] = L; ++out;
                n_next_start[out] = s0 + k; n_next_len[out] = ln - k; n_next_box[out] = R; ++out;
            }
            s_next_nodes = out;
        }
        __syncthreads();
        const int no = s_next_nodes;
        for (int v = tid; v < no; v += DL_LEAF) { n_start[v] = n_next_start[v]; n_len[v] = n_next_len[v]; n_box[v] = n_next_box[v]; }
        if (tid == 0) s_nodes = no;
        DlPoint *t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
    // every level that sorted ended with the swap, and the level that found nothing to sort changed nothing: `cur`
    if (tid < T.len) {
        const DlPoint p = cur[tid];
        const long long i = p.idx;
        const int pos = T.start + tid + sh.shift[T.elem];
        const uint32_t meta = ((uint32_t)molkey[i] << 8) | (uint32_t)T.elem;
        atoms[pos] = make_float4(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], __uint_as_float(meta));
        orig[pos] = (uint32_t)i;
    }
}

struct DlPad { int from[FRMC_MAX_ELEMENTS], to[FRMC_MAX_ELEMENTS]; int nEl; };

__global__ void dl_pad_kernel(DlPad P, float4 *__restrict__ atoms, uint32_t *__restrict__ orig)
{
    const int e = blockIdx.x;
    if (e >= P.nEl) return;
    const float qnan = __int_as_float(0x7FC00000);
    for (int p = P.from[e] + threadIdx.x; p < P.to[e]; p += blockDim.x) {
        atoms[p] = make_float4(qnan, qnan, qnan, __uint_as_float(PAD_META));
        orig[p] = 0xFFFFFFFFu;
    }
}

// ------------------------------------------------------------------ host side
// molecule keys for the meta word (molecule ids when they fit 24 bits, ranks otherwise) and the largest spread of a
// molecule in original indexes (HostLayout::mol_span)
static int molecule_keys(const int32_t *mol, int64_t n, std::vector<int32_t> &rank, const int32_t **keys, uint32_t *span_out)
{
    bool direct = true, runs = true;           // runs: every molecule is one contiguous run of indexes (the usual case)
    int64_t span = 0, run = 0;
    int32_t max_id = -1;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t m = mol[i];
        if (m < 0 || m >= 0x00FFFFFF) direct = false;
        if (m > max_id) max_id = m;
        if (i > 0 && m == mol[i - 1]) { ++run; span = std::max(span, run); }
        else { run = 0; if (i > 0 && m < mol[i - 1]) runs = false; }
    }
    // non-decreasing ids => contiguous runs, and the longest run is the spread; otherwise look every molecule up
    if (!direct) {
        std::vector<int32_t> sorted(mol, mol + n);
        std::sort(sorted.begin(), sorted.end());
        sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
        FRMC_REQUIRE(sorted.size() < 0x00FFFFFFu, FRMC_ELIMIT, "more than 2^24-1 distinct molecules");
        rank.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) rank[(size_t)i] = (int32_t)(std::lower_bound(sorted.begin(), sorted.end(), mol[i]) - sorted.begin());
        max_id = (int32_t)sorted.size() - 1;
        *keys = rank.data();
    } else {
        *keys = mol;
    }
    if (!runs || !direct) {
        const int32_t *k = *keys;
        std::vector<int32_t> first((size_t)max_id + 1, -1);
        span = 0;
        for (int64_t i = 0; i < n; ++i) {
            int32_t &f = first[(size_t)k[i]];
            if (f < 0) f = (int32_t)i; else span = std::max<int64_t>(span, i - f);
        }
    }
    *span_out = (uint32_t)span;
    return FRMC_OK;
}

struct DlTree {
    std::vector<DlTask> split;                 // ordered by level
    std::vector<int> level_start;              // split tasks of level d: [level_start[d], level_start[d+1])
    std::vector<DlTask> leaf;
    int n_boxes = 0;
};

static void dl_walk(DlTree &tree, std::vector<std::vector<DlTask>> &levels, int start, int len, int id, int depth, int elem)
{
    DlTask t;
    memset(&t, 0, sizeof(t));
    t.start = start; t.len = len; t.id = id; t.src = depth & 1; t.elem = elem;
    if (len <= DL_LEAF) { tree.leaf.push_back(t); return; }
    const int unit = 1024;
    t.k = ((len + unit - 1) / unit / 2) * unit;
    t.left = tree.n_boxes++; t.right = tree.n_boxes++;
    if ((int)levels.size() <= depth) levels.resize((size_t)depth + 1);
    levels[(size_t)depth].push_back(t);
    dl_walk(tree, levels, start, t.k, t.left, depth + 1, elem);
    dl_walk(tree, levels, start + t.k, len - t.k, t.right, depth + 1, elem);
}

// Builds the store records of `coords` on device `c` into the context's scratch slots 0 (records) and 1 (original
// indexes), returned through d_atoms_out / d_orig_out.  lay receives what the host keeps: n, npad, nEl, seg_start,
// seg_count, bounds, finiteness, mol_span (rec / orig / inv stay empty); *d_mol_out = the device copy of the molecule
// keys by original index (what the sweep's exact intra/inter test reads).  One stream synchronisation (element counts).
int device_layout(DeviceCtx *c, const float *coords, int64_t n, const int32_t *mol, const int32_t *el, int nEl, int isPBC,
                  HostLayout &lay, float4 **d_atoms_out, uint32_t **d_orig_out, int32_t **d_mol_out, bool raw_on_device)
{
    FRMC_REQUIRE(n >= 0 && n < (1ll << 31) - 4096, FRMC_ELIMIT, "atom count %lld outside 0..2^31", (long long)n);
    FRMC_REQUIRE(nEl >= 1 && nEl <= FRMC_MAX_ELEMENTS, FRMC_ELIMIT, "numberOfElements %d outside 1..%d", nEl, FRMC_MAX_ELEMENTS);
    cudaStream_t st = c->stream;
    const bool timing = getenv("FRMC_LAYOUT_TIMING") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char *what, bool sync) {
        if (!timing) return;
        if (sync) cudaStreamSynchronize(st);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[device layout] %-14s %.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
        tick = now;
    };
    lay.n = n; lay.nEl = nEl;
    lay.seg_count.assign((size_t)nEl, 0);
    lay.seg_start.assign((size_t)nEl + 1, 0);
    lay.rec.clear(); lay.orig.clear(); lay.inv.clear();
    std::vector<int32_t> rank;
    const int32_t *keys = mol;
    lay.mol_span = 0;
    const int n_chunks = (int)((n + DL_CHUNK - 1) / DL_CHUNK);
    // device scratch (slots 7..12 of the context): raw arrays, chunk counts + stats, two point buffers, tasks + boxes
    float *d_coords = (float *)ctx_buffer(c, 7, sizeof(float) * 3 * (size_t)std::max<int64_t>(n, 1));
    int32_t *d_el = (int32_t *)ctx_buffer(c, 8, sizeof(int32_t) * 2 * (size_t)std::max<int64_t>(n, 1));
    int *d_cnt = (int *)ctx_buffer(c, 9, sizeof(int) * FRMC_MAX_ELEMENTS * (size_t)std::max(n_chunks, 1) + sizeof(DlStats));
    if (!d_coords || !d_el || !d_cnt) return FRMC_ENOMEM;
    int32_t *d_key = d_el + std::max<int64_t>(n, 1);
    DlStats *d_stats = reinterpret_cast<DlStats *>(d_cnt + FRMC_MAX_ELEMENTS * (size_t)std::max(n_chunks, 1));
    DlStats h_stats;
    memset(&h_stats, 0, sizeof(h_stats));
    for (int cdim = 0; cdim < 3; ++cdim) { h_stats.lo[cdim] = 0x7FFFFFFF; h_stats.hi[cdim] = (int)0x80000000; }
    FRMC_CUDA(cudaMemcpyAsync(d_stats, &h_stats, sizeof(h_stats), cudaMemcpyHostToDevice, st));
    if (n > 0) {
        // The caller's arrays are pageable: a plain cudaMemcpyAsync stages them through the driver at ~11 GB/s.  Large
        // systems go through the context's page-locked scratch instead, in slices copied by a few host threads, each
        // of which queues the DMA of its own slice as soon as it is staged (20 B/atom: 1.9 -> ~0.8 ms at 10^6 atoms).
        const size_t bytes_c = sizeof(float) * 3 * (size_t)n, bytes_i = sizeof(int32_t) * (size_t)n;
        unsigned char *pin = (n >= 65536 && !raw_on_device) ? (unsigned char *)ctx_pinned(c, bytes_c + 2 * bytes_i) : nullptr;
        if (raw_on_device) {
            // nothing to upload: the arrays are already queued on this stream
        } else if (pin) {
            const int parts = 4;
            std::vector<std::thread> th;
            std::vector<int> errs((size_t)parts, 0);
            const int dev = c->dev;
            for (int t = 0; t < parts; ++t)
                th.emplace_back([&, t] {
                    if (cudaSetDevice(dev) != cudaSuccess) { errs[(size_t)t] = 1; return; }
                    const int64_t a = n * t / parts, b = n * (t + 1) / parts;
                    const size_t cnt = (size_t)(b - a);
                    unsigned char *pc = pin + sizeof(float) * 3 * (size_t)a, *pe = pin + bytes_c + sizeof(int32_t) * (size_t)a,
                                  *pm = pin + bytes_c + bytes_i + sizeof(int32_t) * (size_t)a;
                    memcpy(pc, coords + 3 * a, sizeof(float) * 3 * cnt);
                    if (cudaMemcpyAsync(d_coords + 3 * a, pc, sizeof(float) * 3 * cnt, cudaMemcpyHostToDevice, st) != cudaSuccess) errs[(size_t)t] = 1;
                    memcpy(pe, el + a, sizeof(int32_t) * cnt);
                    if (cudaMemcpyAsync(d_el + a, pe, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st) != cudaSuccess) errs[(size_t)t] = 1;
                    memcpy(pm, mol + a, sizeof(int32_t) * cnt);
                    if (cudaMemcpyAsync(d_key + a, pm, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st) != cudaSuccess) errs[(size_t)t] = 1;
                });
            for (auto &t : th) t.join();
            for (int e : errs) FRMC_REQUIRE(!e, FRMC_ECUDA, "staged upload of the atom arrays failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else {
            FRMC_CUDA(cudaMemcpyAsync(d_coords, coords, bytes_c, cudaMemcpyHostToDevice, st));
            FRMC_CUDA(cudaMemcpyAsync(d_el, el, bytes_i, cudaMemcpyHostToDevice, st));
            FRMC_CUDA(cudaMemcpyAsync(d_key, mol, bytes_i, cudaMemcpyHostToDevice, st));
        }
        dl_count_kernel<<<n_chunks, 256, 0, st>>>(d_coords, d_el, d_key, (long long)n, nEl, d_cnt, d_stats);
        FRMC_LAUNCH_CHECK();
        dl_scan_kernel<<<1, 1024, 0, st>>>(d_cnt, n_chunks, nEl, d_stats);
        FRMC_LAUNCH_CHECK();
    }
    FRMC_CUDA(cudaMemcpyAsync(&h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
    FRMC_CUDA(cudaStreamSynchronize(st));
    lap("upload+count", false);
    FRMC_REQUIRE(!h_stats.bad_element, FRMC_EINVAL, "an elementIndex entry lies outside 0..%d", nEl - 1);
    // molecule keys of the meta word and the spread of a molecule: nothing to do on the host for the usual atomic
    // system (every atom its own molecule, ids in range: the device saw no two neighbours sharing one and the ids
    // never decreasing, so no molecule has two atoms); otherwise one pass over the ids (and a ranking when they do
    // not fit 24 bits), with the ranked keys replacing the raw ids on the device
    if (n > 0 && (h_stats.mol_not_direct || h_stats.mol_decreasing || h_stats.mol_repeats)) {
        int rc = molecule_keys(mol, n, rank, &keys, &lay.mol_span);
        if (rc) return rc;
        if (keys != mol) FRMC_CUDA(cudaMemcpyAsync(d_key, keys, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    }
    *d_mol_out = d_key;
    lap("molecule keys", false);
    DlElemBase eb;
    DlShift sh;
    DlPad pad;
    memset(&eb, 0, sizeof(eb)); memset(&sh, 0, sizeof(sh)); memset(&pad, 0, sizeof(pad));
    pad.nEl = nEl;
    int64_t compact = 0;
    for (int e = 0; e < nEl; ++e) {
        lay.seg_count[(size_t)e] = h_stats.count[e];
        const int64_t padded = (lay.seg_count[(size_t)e] + SEG_PAD - 1) / SEG_PAD * SEG_PAD;
        lay.seg_start[(size_t)e + 1] = lay.seg_start[(size_t)e] + padded;
        eb.base[e] = (int)compact;
        sh.shift[e] = (int)(lay.seg_start[(size_t)e] - compact);
        pad.from[e] = (int)(lay.seg_start[(size_t)e] + lay.seg_count[(size_t)e]);
        pad.to[e] = (int)lay.seg_start[(size_t)e + 1];
        compact += lay.seg_count[(size_t)e];
    }
    lay.npad = lay.seg_start[(size_t)nEl];
    FRMC_REQUIRE(lay.npad < (1ll << 31), FRMC_ELIMIT, "padded atom count exceeds 2^31");
    lay.finite = !h_stats.not_finite;
    for (int cdim = 0; cdim < 3; ++cdim) {
        lay.lo[cdim] = (n > 0 && h_stats.lo[cdim] <= h_stats.hi[cdim]) ? dl_ordered_to_float(h_stats.lo[cdim]) : 0.f;
        lay.hi[cdim] = (n > 0 && h_stats.lo[cdim] <= h_stats.hi[cdim]) ? dl_ordered_to_float(h_stats.hi[cdim]) : 0.f;
    }
    if (!lay.finite) lay.hi[0] = INFINITY;   // forces the general wrap (build_layout does the same)

    float4 *d_atoms = (float4 *)ctx_buffer(c, 0, sizeof(float4) * (size_t)std::max<int64_t>(lay.npad, 1));
    uint32_t *d_orig = (uint32_t *)ctx_buffer(c, 1, sizeof(uint32_t) * (size_t)std::max<int64_t>(lay.npad, 1));
    if (!d_atoms || !d_orig) return FRMC_ENOMEM;
    *d_atoms_out = d_atoms; *d_orig_out = d_orig;
    if (n == 0) return FRMC_OK;

    // the tree shape (host) and the root boxes
    DlTree tree;
    std::vector<std::vector<DlTask>> levels;
    std::vector<DlBox> roots;
    std::vector<int> root_ids;
    for (int e = 0; e < nEl; ++e) {
        if (lay.seg_count[(size_t)e] == 0) continue;
        DlBox b;
        for (int cdim = 0; cdim < 3; ++cdim) {
            b.lo[cdim] = isPBC ? 0.f : (lay.finite ? lay.lo[cdim] : 0.f);
            b.hi[cdim] = isPBC ? 1.f : (lay.finite ? lay.hi[cdim] : 0.f);
        }
        const int id = tree.n_boxes++;
        roots.push_back(b); root_ids.push_back(id);
        dl_walk(tree, levels, eb.base[e], (int)lay.seg_count[(size_t)e], id, 0, e);
    }
    std::vector<DlBox> h_boxes((size_t)tree.n_boxes);
    for (size_t r = 0; r < roots.size(); ++r) h_boxes[(size_t)root_ids[r]] = roots[r];
    // large nodes first inside a level: they go through the multi-CTA split (dl_whist / dl_wcut / dl_wscatter)
    static const bool wide_on = []() { const char *e = getenv("FRMC_WIDE_SPLIT"); return !(e && atoi(e) == 0); }();
    std::vector<int> level_wide;                   // wide tasks of level d
    std::vector<DlWChunk> wchunks;                 // all levels, level after level
    std::vector<int> wchunk_first;                 // per wide task (+1 per level): first chunk, relative to the level's first chunk
    std::vector<int> level_chunk0, level_first0;   // per level: offsets into wchunks / wchunk_first
    size_t max_level_wide = 0;
    tree.level_start.push_back(0);
    for (auto &lv : levels) {
        std::stable_partition(lv.begin(), lv.end(), [](const DlTask &t) { return wide_on && t.len >= DL_WIDE_MIN; });
        int nw = 0;
        while (nw < (int)lv.size() && wide_on && lv[(size_t)nw].len >= DL_WIDE_MIN) ++nw;
        level_wide.push_back(nw);
        level_chunk0.push_back((int)wchunks.size());
        level_first0.push_back((int)wchunk_first.size());
        const int c_base = (int)wchunks.size();
        for (int t = 0; t < nw; ++t) {
            wchunk_first.push_back((int)wchunks.size() - c_base);
            for (int first = 0; first < lv[(size_t)t].len; first += DL_WCHUNK) {
                DlWChunk ch;
                ch.task = t; ch.first = first; ch.len = std::min(DL_WCHUNK, lv[(size_t)t].len - first);
                wchunks.push_back(ch);
            }
        }
        wchunk_first.push_back((int)wchunks.size() - c_base);
        max_level_wide = std::max(max_level_wide, (size_t)nw);
        tree.split.insert(tree.split.end(), lv.begin(), lv.end());
        tree.level_start.push_back((int)tree.split.size());
    }
    size_t max_level_chunks = 0;
    for (size_t d = 0; d < levels.size(); ++d)
        max_level_chunks = std::max(max_level_chunks, (size_t)((d + 1 < levels.size() ? level_chunk0[d + 1] : (int)wchunks.size()) - level_chunk0[d]));
    const size_t n_tasks = tree.split.size() + tree.leaf.size();
    DlPoint *d_pts = (DlPoint *)ctx_buffer(c, 10, sizeof(DlPoint) * 2 * (size_t)n);
    unsigned char *d_tb = (unsigned char *)ctx_buffer(c, 11, sizeof(DlTask) * n_tasks + sizeof(DlBox) * (size_t)tree.n_boxes + 64);
    if (!d_pts || !d_tb) return FRMC_ENOMEM;
    DlPoint *buf0 = d_pts, *buf1 = d_pts + n;
    DlTask *d_tasks = reinterpret_cast<DlTask *>(d_tb);
    DlBox *d_boxes = reinterpret_cast<DlBox *>(d_tb + sizeof(DlTask) * n_tasks);
    std::vector<DlTask> all(tree.split);
    all.insert(all.end(), tree.leaf.begin(), tree.leaf.end());
    FRMC_CUDA(cudaMemcpyAsync(d_tasks, all.data(), sizeof(DlTask) * n_tasks, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_boxes, h_boxes.data(), sizeof(DlBox) * h_boxes.size(), cudaMemcpyHostToDevice, st));
    // scratch of the multi-CTA splits (slot 12): chunk descriptors and first-chunk tables of every level, then per level
    // (reused) the chunks' histograms, the nodes' histograms, the cuts and the chunk offsets
    DlWChunk *d_wchunks = nullptr;
    int *d_wfirst = nullptr, *d_chunk_hist = nullptr, *d_node_hist = nullptr;
    DlWCut *d_cuts = nullptr;
    int3 *d_chunk_off = nullptr;
    if (!wchunks.empty()) {
        auto up16 = [](size_t b) { return (b + 15) / 16 * 16; };
        const size_t b_chunks = up16(sizeof(DlWChunk) * wchunks.size()), b_first = up16(sizeof(int) * wchunk_first.size()),
                     b_chist = up16(sizeof(int) * DL_BINS * max_level_chunks), b_nhist = up16(sizeof(int) * DL_BINS * max_level_wide),
                     b_cuts = up16(sizeof(DlWCut) * max_level_wide), b_off = up16(sizeof(int3) * max_level_chunks);
        unsigned char *d_w = (unsigned char *)ctx_buffer(c, 12, b_chunks + b_first + b_chist + b_nhist + b_cuts + b_off);
        if (!d_w) return FRMC_ENOMEM;
        d_wchunks = reinterpret_cast<DlWChunk *>(d_w);
        d_wfirst = reinterpret_cast<int *>(d_w + b_chunks);
        d_chunk_hist = reinterpret_cast<int *>(d_w + b_chunks + b_first);
        d_node_hist = reinterpret_cast<int *>(d_w + b_chunks + b_first + b_chist);
        d_cuts = reinterpret_cast<DlWCut *>(d_w + b_chunks + b_first + b_chist + b_nhist);
        d_chunk_off = reinterpret_cast<int3 *>(d_w + b_chunks + b_first + b_chist + b_nhist + b_cuts);
        FRMC_CUDA(cudaMemcpyAsync(d_wchunks, wchunks.data(), sizeof(DlWChunk) * wchunks.size(), cudaMemcpyHostToDevice, st));
        FRMC_CUDA(cudaMemcpyAsync(d_wfirst, wchunk_first.data(), sizeof(int) * wchunk_first.size(), cudaMemcpyHostToDevice, st));
        FRMC_CUDA(cudaMemsetAsync(d_node_hist, 0, b_nhist, st));          // dl_wcut_kernel re-arms it level after level
    }
    lap("tree+tasks", false);
    dl_gather_kernel<<<n_chunks, 256, 0, st>>>(d_coords, d_el, (long long)n, nEl, isPBC ? 1 : 0, d_cnt, eb, buf0);
    FRMC_LAUNCH_CHECK();
    for (size_t d = 0; d + 1 < tree.level_start.size(); ++d) {
        const int a = tree.level_start[d], b = tree.level_start[d + 1];
        const int nw = level_wide[d];
        if (nw > 0) {
            const int ch0 = level_chunk0[d];
            const int nch = (d + 1 < level_chunk0.size() ? level_chunk0[d + 1] : (int)wchunks.size()) - ch0;
            dl_whist_kernel<<<nch, 1024, 0, st>>>(d_tasks + a, d_wchunks + ch0, buf0, buf1, d_boxes, d_chunk_hist, d_node_hist);
            FRMC_LAUNCH_CHECK();
            dl_wcut_kernel<<<nw, 1024, 0, st>>>(d_tasks + a, d_wfirst + level_first0[d], d_chunk_hist, d_node_hist, d_boxes, d_cuts, d_chunk_off,
                                                d_wchunks + ch0);
            FRMC_LAUNCH_CHECK();
            dl_wscatter_kernel<<<nch, 1024, 0, st>>>(d_tasks + a, d_wchunks + ch0, d_cuts, d_chunk_off, buf0, buf1);
            FRMC_LAUNCH_CHECK();
        }
        if (b > a + nw) {
            dl_split_kernel<<<b - a - nw, 1024, 0, st>>>(d_tasks + a + nw, buf0, buf1, d_boxes);
            FRMC_LAUNCH_CHECK();
        }
    }
    lap("gather+splits", true);
    static const bool leaf_fast_on = []() { const char *e = getenv("FRMC_LEAF_FAST"); return !(e && atoi(e) == 0); }();
    dl_leaf_kernel<<<(unsigned)tree.leaf.size(), DL_LEAF, 0, st>>>(d_tasks + tree.split.size(), buf0, buf1, d_boxes, d_coords, d_key, sh, d_atoms, d_orig,
                                                                   (leaf_fast_on && lay.finite) ? 1 : 0);
    FRMC_LAUNCH_CHECK();
    dl_pad_kernel<<<nEl, 256, 0, st>>>(pad, d_atoms, d_orig);
    FRMC_LAUNCH_CHECK();
    lap("leaves+pad", true);
    return FRMC_OK;
}

}  // namespace frmc
