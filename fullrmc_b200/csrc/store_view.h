// store_view.h -- what another translation unit may see of a device store (store.cu owns the struct).
#pragma once
#include "common.cuh"

#include <vector>

struct frmc_store;

namespace frmc {

static const int FRMC_MAX_GROUP_ = FRMC_MAX_GROUP;

struct ProposalIn {                  // passed BY VALUE as a kernel parameter (constant bank)
    int k;
    int pos[FRMC_MAX_GROUP];         // positions in the sorted store (host lookup in the inverse permutation)
    float moved[3 * FRMC_MAX_GROUP]; // moved box coordinates
};

struct Proposal {                    // device copy kept for the commit kernel
    int k;
    int pos[FRMC_MAX_GROUP];
    float4 newc[FRMC_MAX_GROUP];     // same meta, moved coordinates
};

// Read-only view of the store for kernels that evaluate something else against the same atoms (the distance
// constraints, atomdist.cu).  `pending` = 1 means the last proposal was accepted but the device has not applied it
// yet (the fused path defers the commit to the next launch): the records at prop->pos[0..k) are then to be read
// as prop->newc[].  Nothing is flushed, so the deferred-commit fast path of the histogram constraints survives.
struct StoreView {
    int dev;
    cudaStream_t stream;
    int sm_count;
    const float4 *atoms;
    const uint32_t *orig;
    int64_t n, npad;                 // n: atoms the store holds now
    int64_t n0;                      // atoms of the layout (the numbering of inv / orig)
    uint64_t layout_gen;             // changes whenever the store is laid out again (positions of the records change)
    const int32_t *rel2real;         // host: the engine's relative index -> original index once atoms were removed (NULL: identity)
    const int32_t *inv;              // host: original index -> position
    Lattice L;
    int isPBC;
    float lo[3], hi[3];              // coordinate bounds seen so far (wrap-mode choice)
    int pending;                     // 0 nothing, 1 accept pending, 2 reject pending
    const Proposal *prop;            // device
    DeviceCtx *ctx;
};

int store_view(frmc_store *s, StoreView *out);   // store.cu; ends a persistent run (its kernel owns the GPU)
int store_flush(frmc_store *s);                  // store.cu; applies a deferred accept / reject to the device state
void storedist_release(frmc_store *s);           // storedist.cu; frees the distance constraints registered on a store
void storecoord_release(frmc_store *s);          // storecoord.cu; frees the coordination-number constraints registered on a store

}  // namespace frmc
