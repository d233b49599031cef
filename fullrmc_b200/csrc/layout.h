// layout.h -- the device-resident atom store layout (host-side description).
//
// Atoms are kept ELEMENT-SORTED and, inside an element, in K-D ORDER of the (periodically
// reduced) box coordinates -- every aligned run of 1024 / 256 / 32 records is a compact box --
// in one array of 16-byte records
//     float4 { x, y, z, meta }      meta = (molecule_rank << 8) | element      (u32 bits)
// plus a parallel u32 array `orig` with the atom's original index.  Each element's
// segment is padded to a multiple of SEG_PAD records with NaN coordinates
// (meta = 0xFFFFFFFF, orig = 0xFFFFFFFF): NaN fails every range test, so padding
// never contributes and no kernel needs a bounds check inside a segment.
//
// Why k-d order: a block of SEG_PAD consecutive records is then spatially compact, its
// bounding box is small, and the full-histogram kernel skips every (I block, J block) whose boxes
// are provably farther apart than maxDistance (fullhist.cu).  Counts are integers, so the order
// in which pairs are visited never shows in the result.
//
// Why element-sorted: a (tile I, tile J) pair then touches one unordered element pair, so the
// full-histogram kernel needs only 4 x histSize shared-memory counters per CTA
// ({intra,inter} x {[a,b],[b,a]}) instead of 2 x nEl^2 x histSize, whatever nEl is.
// The reference's ORDERED output ([el[i], el[j]] with i<j in original order,
// pairs_histograms.pyx:289-335) is recovered from `orig`.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace frmc {

static const int SEG_PAD = 256;
static const uint32_t PAD_META = 0xFFFFFFFFu;

struct HostLayout {
    int64_t n = 0;           // real atoms
    int64_t npad = 0;        // records including padding
    int nEl = 0;
    std::vector<int64_t> seg_start;   // [nEl+1] padded start position of each element segment
    std::vector<int64_t> seg_count;   // [nEl]   real atoms per element
    std::vector<float> rec;           // [npad*4] x,y,z,meta(bits)
    std::vector<uint32_t> orig;       // [npad]  original index or 0xFFFFFFFF
    std::vector<int32_t> inv;         // [n]     original index -> position
    std::vector<float> kd;            // scratch of the k-d ordering (4 floats per atom), kept to reuse its pages
    float lo[3], hi[3];               // coordinate bounds (mode selection)
    bool finite = true;
    uint32_t mol_span = 0;            // largest |i - j| (original indexes) over pairs of atoms of one molecule
};

// Builds the sorted layout.  Returns 0 or a negative FRMC_E* code (error string set).
// isPBC selects how a coordinate maps to a cell (fractional part vs. position inside the bounds).
int build_layout(const float *coords, int64_t n, const int32_t *mol, const int32_t *el, int nEl, int isPBC, HostLayout &out);

// The same layout built on the device from the caller's host arrays (devlayout.cu): records and original indexes land in
// the context's scratch slots 0 / 1; `lay` receives the host-side description only (no rec / orig / inv).
struct DeviceCtx;
int device_layout(DeviceCtx *c, const float *coords, int64_t n, const int32_t *mol, const int32_t *el, int nEl, int isPBC,
                  HostLayout &lay, float4 **d_atoms_out, uint32_t **d_orig_out, int32_t **d_mol_out, bool raw_on_device = false);
// raw_on_device: the caller has already queued the raw arrays into the context's slots 7 (coords [3n] floats) and 8
// (el [n] | mol [n] ints) on the context's stream (multigpu.cu: every GPU brings its slice over its own PCIe link)

// One ROW of full-histogram work: I-tile [i0, i0 + 256*ni) x the J range [j0, j1) of one element pair
// (padded positions).  ea/eb are the (segment) elements of the two ranges; when ea == eb the J range starts
// at the tile and only pairs p<q count.  The device cuts rows into items of surviving blocks.
struct WorkItem {
    int32_t i0, ni, j0, j1;
    int32_t ea, eb, pad0, pad1;
};

// Builds the upper-triangle row list, ordered by element pair (a CTA walking it flushes its
// shared-memory counters only when the pair changes); rows with index % nshards == shard are kept.
void build_rows(const HostLayout &lay, int R, int shard, int nshards, std::vector<WorkItem> &rows);

// Device-side lists of surviving block pairs (fullhist.cu), grow-only, owned by whoever launches the kernel.
struct PairLists {
    int *row_ints = nullptr;        // [4*n_rows + 2] count / items / start / item start per row + two totals
    uint32_t *entries = nullptr;    // surviving J block of every (row, k)
    int4 *items = nullptr;          // {row, first entry, blocks, -}
    size_t row_cap = 0, entries_cap = 0, items_cap = 0;
    int n_entries = 0, n_items = 0; // of the last launch
    int n_rows = 0;                 // rows of the last launch (row_ints + 3 * n_rows = first item of every row)
    int item_blocks = 8;            // surviving blocks per item
    int *pair_next = nullptr;       // full histogram: one task counter per element pair (+ scratch), zeroed per launch
    float2 *bin_table = nullptr;    // full histogram: (T[b], T[b+1]) d^2 thresholds of the bin edges
    size_t bin_cap = 0;
    float4 *recs = nullptr;         // full histogram: sweep records {x, y, z, original index}
    size_t recs_cap = 0;
    void release();
};

struct CullParams;
// box pass + surviving block pairs of `rows`, cut into items (fullhist.cu)
int build_pair_lists(cudaStream_t stream, const float4 *atoms, int64_t npad, float4 *bbox, const WorkItem *rows, int n_rows,
                     const CullParams &cp, PairLists &lists);

}  // namespace frmc
