/*
 * fullrmc_b200.h -- C ABI of the B200-native pair-histogram backend for fullrmc.
 *
 * This is the drop-in boundary.  Every entry point takes
 * plain pointers and sizes (no torch / numpy / C++ types), returns 0 on success
 * and a negative FRMC_E* code on failure; the message is available from
 * frmc_last_error().  No exception crosses this boundary.  Handles are not
 * thread-safe (one host thread per store, like the reference Engine).
 *
 * Citations are relative to the reference tree (bachiraoun/fullrmc v4.1.0).
 *
 * Arithmetic contract: distances and bin indices are computed in fp32 with the
 * reference's exact operation order, no FMA contraction, IEEE sqrt and divide,
 * round-half-away-from-zero minimum image (Extensions/pairs_distances.pyx:31-32),
 * bin rule `d<min skip; d>=max skip; (int)((d-min)/bin)`
 * (Extensions/pairs_histograms.pyx:58-63).  Counts are integers on the device
 * and converted to float32 at this boundary (exact below 2^24 per cell, where the
 * reference's own `+= 1.0f` saturates).  A bin index that rounds up to histSize
 * (undefined behaviour in the reference, boundscheck(False)) is dropped and
 * counted in *edge_overflow.
 */
#ifndef FULLRMC_B200_H
#define FULLRMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRMC_OK 0
#define FRMC_EINVAL (-1)    /* bad argument (shape, range, NULL) */
#define FRMC_ECUDA (-2)     /* CUDA runtime error */
#define FRMC_ENOMEM (-3)    /* allocation failure */
#define FRMC_ESTATE (-4)    /* call not valid in the handle's current state */
#define FRMC_ELIMIT (-5)    /* exceeds a documented limit (elements, molecules, group size) */

#define FRMC_MAX_ELEMENTS 16      /* distinct element indices per store */
#define FRMC_MAX_GROUP 64         /* atoms moved by one proposal */
#define FRMC_MAX_GRIDS 4          /* r-grids per store */
#define FRMC_MAX_MODELS 8         /* models (constraints) per store */

/* model kinds: which constraint-level total is produced from the histogram */
#define FRMC_KIND_PDF 0   /* G(r)   PairDistributionConstraints.py:847-895   */
#define FRMC_KIND_PCF 1   /* g(r)   PairCorrelationConstraints.py:126-169    */
#define FRMC_KIND_SQ 2    /* S(Q)   StructureFactorConstraints.py:780-822    */
#define FRMC_KIND_RSQ 3   /* S(Q)-1 StructureFactorConstraints.py:1253-1260  */

const char *frmc_last_error(void);
const char *frmc_version(void);
/* Edge-bin policy for grids / stateless calls created AFTER the call.  The reference writes
 * hist[a,b,bin] with boundscheck(False); when fp32 rounding makes bin == histSize (a pair within an
 * ulp of maxDistance; happens in Examples/atomicNiTi within 10 moves) that write lands in the next
 * [a,b] slab's first bin (or outside the array for the last slab).  on=0 (default): such events are
 * dropped; on=1: the in-array spill is reproduced so trajectories stay bit-identical to the
 * reference's.  Either way they are counted in edge_overflow.  Returns the previous setting. */
int frmc_set_edge_spill(int on);
/* Block culling of the full-histogram kernels (default on): atoms are stored element by element in k-d order,
 * each 256-atom block and each 32-atom sub-block carries a bounding box, and block pairs provably farther apart
 * than maxDistance are not swept.  The result is identical either way (tests sweep both); on=0 forces the
 * reference's plain O(N^2) sweep, for measurement.  Returns the previous setting. */
int frmc_set_block_culling(int on);
/* The finer level of that culling (default on): inside every 32-record unit the sweep also skips the 8-record chunks whose
 * own boxes are out of reach of the warp's 32 atoms (csrc/fullhist.cu: sweep_records_kernel).  on=0 keeps the unit-level
 * culling only, for measurement (frmc_store_swept_pairs then counts whole units).  Identical results either way. */
int frmc_set_chunk_culling(int on);
/* Where the stateless full histogram orders the caller's atoms (default on = on the device, csrc/devlayout.cu:
 * raw arrays are uploaded as they are; off = k-d ordering on the host cores before the upload).  The histogram is
 * identical either way.  Returns the previous setting. */
int frmc_set_device_layout(int on);
/* Device time of the dominant kernel of the stateless entry points (full histogram: box pass + lists + sweep; distance
 * windows: the block sweep; coordination numbers: the counting kernel): CUDA events around it on the library's stream
 * when switched on; frmc_ctx_kernel_ms returns the last call's figure (for bench.py's rooflines). */
int frmc_ctx_set_timing(int dev, int on);
int frmc_ctx_kernel_ms(int dev, double *ms);
/* number of visible CUDA devices, or a negative error code (no CPU fallback exists) */
int frmc_device_count(void);

/* ------------------------------------------------------------------------------------
 * Stateless entry points: one per reference extension function on the hot path.
 * Inputs are HOST pointers, borrowed for the duration of the call; outputs are HOST
 * buffers owned by the caller.  coords are C-contiguous [n,3] fp32 (fractional when
 * isPBC, Cartesian otherwise); basis is [3,3] row-major, rows = lattice vectors.
 * ---------------------------------------------------------------------------------- */

/* Generic "k points against n coords" kernel behind the ten functions of
 * Extensions/pairs_distances.pyx (def wrappers :483-1024, cdef kernels :45-471).
 *   points     [k,3]                 explicit points (from_index == NULL) or ignored
 *   from_index [k] or NULL           take point t = coords[from_index[t]]
 *   start      [k] or NULL           first coords row written for point t (allAtoms=False -> index)
 *   ibc_sign   +1: point-coords[i] (every difference kernel of the reference, :142-270)
 *              -1: coords[i]-point (sign used inside the IBC to-point distance kernel :414-434;
 *                  only visible when want_diff=1)
 *   want_diff  0: distances, out is [n,k]; 1: difference vectors, out is [n,3,k]
 * Rows below start[t] are written as 0 (the reference leaves them uninitialised). */
int frmc_points_to_coords(int dev, const float *points, const int32_t *from_index, const int64_t *start,
                          int64_t k, const float *coords, int64_t n, const float *basis, int isPBC,
                          int ibc_sign, int want_diff, float *out);

/* Extensions/pairs_distances.pyx:483-525 from_to_points_differences: row-wise
 * boundaryConditions(pointsTo[i]-pointsFrom[i]); out is [n,3]. */
int frmc_from_to_points_differences(int dev, const float *points_from, const float *points_to, int64_t n,
                                    const float *basis, int isPBC, float *out);

/* Extensions/boundary_conditions_collection.pyx:88-110 transform_coordinates: out [n,3] = coords [n,3] . transMatrix [3,3]
 * (row vectors; every product and sum rounded to float32 in the reference's order).  Engine.py:3223 maps moved real
 * coordinates to box coordinates with it (transMatrix = reciprocalBasisVectors). */
int frmc_transform_coordinates(int dev, const float *trans_matrix, const float *coords, int64_t n, float *out);

/* Extensions/pairs_histograms.pyx:150-217 multiple_pairs_histograms_coords.
 * hintra/hinter: [nEl,nEl,hs] fp32, overwritten (the reference returns fresh np.zeros arrays). */
int frmc_multiple_pairs_histograms_coords(int dev, const int32_t *indexes, int64_t k, const float *coords,
                                          int64_t n, const float *basis, int isPBC, const int32_t *mol,
                                          const int32_t *el, int nEl, float rmin, float rmax, float bin,
                                          int hs, int allAtoms, float *hintra, float *hinter,
                                          uint64_t *edge_overflow);

/* Extensions/pairs_histograms.pyx:289-335 full_pairs_histograms_coords (ordered upper
 * triangle [el[i],el[j]], i<j).  shard/nshards split the tile work list across callers
 * (one process per GPU); each caller gets its partial histogram, the sum over shards is
 * the full result.  nshards=1 for the plain drop-in call. */
int frmc_full_pairs_histograms_coords(int dev, const float *coords, int64_t n, const float *basis, int isPBC,
                                      const int32_t *mol, const int32_t *el, int nEl, float rmin,
                                      float rmax, float bin, int hs, int shard, int nshards,
                                      float *hintra, float *hinter, uint64_t *edge_overflow);

/* The same histogram over `ndev` GPUs of one box from ONE call of one process ("shards + NCCL
 * inside"; what an unmodified Engine's compute_data reaches through fullrmc_b200.Core.pairs_histograms when
 * $FULLRMC_B200_DEVICES lists several devices): devs[0] orders the atoms and copies the store to the others over
 * NVLink, every device sweeps its share of the triangular row list, one ncclAllReduce(sum) of the 64-bit counts
 * (in-process communicator, ncclCommInitAll once per device set; libnccl.so.2 is loaded at run time,
 * $FULLRMC_B200_NCCL overrides its path), devs[0] returns the arrays.  Identical to the one-device result. */
int frmc_full_pairs_histograms_coords_multi(int ndev, const int *devs, const float *coords, int64_t n, const float *basis,
                                            int isPBC, const int32_t *mol, const int32_t *el, int nEl, float rmin,
                                            float rmax, float bin, int hs, float *hintra, float *hinter,
                                            uint64_t *edge_overflow);
/* how the last multi-device call combined the devices' counts: "nccl <version> x<ndev>" or "single device" */
const char *frmc_multi_reduce_path(void);

/* Host-only inspection of the store layout (no device needed): original index of every record of the
 * element-sorted, k-d ordered store (0xFFFFFFFF = padding), padded record count, padded segment start of
 * every element (nEl + 1 values).  orig_out must hold n + 256 * nEl records. */
int frmc_debug_layout(int64_t n, const float *coords, const int32_t *mol, const int32_t *el, int nEl, int isPBC,
                      int64_t capacity, uint32_t *orig_out, int64_t *npad_out, int64_t *seg_start_out);
/* The same inspection of the layout the DEVICE builds for the stateless full histogram (csrc/devlayout.cu). */
int frmc_debug_device_layout(int dev, int64_t n, const float *coords, const int32_t *mol, const int32_t *el, int nEl, int isPBC,
                             int64_t capacity, uint32_t *orig_out, int64_t *npad_out, int64_t *seg_start_out);
/* Host-only view of the multi-GPU decomposition of the full histogram (runs without a device):
 * work items and atom pairs assigned to `shard` of `nshards`; over all shards the pairs sum to n(n-1)/2. */
int frmc_debug_work_items(int64_t n, const int32_t *el, int nEl, int shard, int nshards, int sm_count,
                          int64_t *n_items, int64_t *n_pairs);

/* Extensions/pairs_histograms.pyx:225-281 multiple_pairs_histograms_dists; distances is
 * [n,k] row-major, column t belongs to indexes[t].  (:343-383 full_* = arange, allAtoms=0) */
int frmc_multiple_pairs_histograms_dists(int dev, const int32_t *indexes, int64_t k, const float *distances,
                                         int64_t n, const int32_t *mol, const int32_t *el, int nEl,
                                         float rmin, float rmax, float bin, int hs, int allAtoms,
                                         float *hintra, float *hinter, uint64_t *edge_overflow);

/* Extensions/pairs_histograms.pyx:77-141 single_pairs_histograms: IN-PLACE update of
 * hintra/hinter from one precomputed distance row (element stride dstride). */
int frmc_single_pairs_histograms(int dev, int32_t atomIndex, const float *distances, int64_t dstride,
                                 int64_t n, const int32_t *mol, const int32_t *el, int nEl, int hs,
                                 float *hintra, float *hinter, float rmin, float rmax, float bin,
                                 int allAtoms, uint64_t *edge_overflow);

/* Extensions/reciprocal_space.pyx:82-109 Gr_to_sq and :42-73 gr_to_sq (double-precision
 * term, fp32 accumulate, r in index order) and :118-145 sq_to_Gr (documented math; the
 * reference function itself raises TypeError, so it has no oracle). */
int frmc_Gr_to_sq(int dev, const float *distances, const float *Gr, int64_t n, const float *qrange, int64_t m,
                  float *sq);
int frmc_gr_to_sq(int dev, const float *distances, const float *gr, int64_t n, const float *qrange, int64_t m,
                  float rho, float *sq);
int frmc_sq_to_Gr(int dev, const float *qvalues, const float *rvalues, const float *sq, int64_t m, int64_t n,
                  float *Gr);

/* Shape function of a finite system on the device (ShapeFunction.get_Gr_shape_function,
 * Constraints/Collection.py:83-110, and the private StructureFactorConstraint's total behind it): from the ordered
 * histograms hintra / hinter [nEl,nEl,hs] of the whole system on the shape grid (shell volumes / centres [hs]) to
 * G_shape on the r values r[nr], through S(q)-1 on q[nq].  pair_a / pair_b / pair_coef [n_pairs]: the unordered
 * element pairs and w_ij / D_ij of each (weighting scheme over N_ij / volume).  Double-precision sums; within 1e-6
 * (norm-wise) of the reference's float32 numpy path. */
int frmc_shape_function(int dev, const float *hintra, const float *hinter, int nEl, int hs, int n_pairs, const int32_t *pair_a,
                        const int32_t *pair_b, const double *pair_coef, const float *shell_volumes, const float *shell_centers,
                        double rho0, const float *q, int nq, const float *r, int nr, float *out);

/* ---- distance-constraint kernels (Extensions/atomic_distances.pyx) --------------
 * multiple_atomic_distances_coords (:326-417) / full_atomic_distances_coords (:500-567): counts and float32 sums
 * of the (optionally reduced) distances of the pairs inside -- or, without FRMC_AD_WITHIN, outside -- the
 * [lower, upper) window of their type pair.  lowerLimit / upperLimit [nT*nT] are indexed [type_i, type_a];
 * nintra, dintra, ninter, dinter [nT*nT] are indexed [type_a, type_i] (the reference's [nT,nT,1] arrays).
 * The sums are accumulated in the reference's loop order, so they are bit-identical to it. */
#define FRMC_AD_INTER 1      /* interMolecular        */
#define FRMC_AD_INTRA 2      /* intraMolecular        */
#define FRMC_AD_WITHIN 4     /* countWithinLimits     */
#define FRMC_AD_TO_UPPER 8   /* reduceDistanceToUpper */
#define FRMC_AD_TO_LOWER 16  /* reduceDistanceToLower */
#define FRMC_AD_REDUCE 32    /* reduceDistance        */
int frmc_multiple_atomic_distances_coords(int dev, const int32_t *indexes, int64_t k, const float *coords, int64_t n,
                                          const float *basis, int isPBC, const int32_t *mol, const int32_t *type, int nT,
                                          const float *lowerLimit, const float *upperLimit, int flags, int allAtoms,
                                          int32_t *nintra, float *dintra, int32_t *ninter, float *dinter);
int frmc_full_atomic_distances_coords(int dev, const float *coords, int64_t n, const float *basis, int isPBC,
                                      const int32_t *mol, const int32_t *type, int nT, const float *lowerLimit,
                                      const float *upperLimit, int flags, int32_t *nintra, float *dintra,
                                      int32_t *ninter, float *dinter);

/* ---- coordination-number counts (Extensions/atomic_coordination.pyx) ----------
 * One call = a flat list of tasks; task t counts the atoms j of list task_list[t] whose distance to atom
 * task_core[t] satisfies task_lower[t] <= d <= task_upper[t] (both ends inclusive, the float32 distance of
 * pairs_distances_to_point; atomic_coordination.pyx:31-48, :89-108) and adds the count to counts[task_out[t]].
 * Lists are given as offsets [nlists+1] into list_indexes.  This is what single_atom_single_shell_coords (:89),
 * single_atom_multi_shells_coords (:139), single_atom_coord_number_coords (:207), multi_atoms_coord_number_coords
 * (:280) and all_atoms_coord_number_coords (:349) reduce to; the *_totdists forms (:71, :116, :171, :249, :317)
 * pass `distances` [nrows, n] instead of coords (coords NULL) and task_core[t] is then the ROW of the matrix.
 * counts [nout] is overwritten (int32; the reference accumulates the same integers as float32). */
int frmc_coordination_counts(int dev, const float *coords, int64_t n, const float *basis, int isPBC,
                             const float *distances, int64_t nrows, int64_t ntasks, const int32_t *task_core,
                             const int32_t *task_list, const int32_t *task_out, const float *task_lower,
                             const float *task_upper, int64_t nlists, const int64_t *list_offsets,
                             const int32_t *list_indexes, int64_t nout, int32_t *counts);

/* ------------------------------------------------------------------------------------
 * Stateful fast path: device-resident coordinate store + running histograms.
 * Replaces compute_data / compute_before_move / compute_after_move / accept_move /
 * reject_move of the three constraints (PairDistributionConstraints.py:1001-1166,
 * PairCorrelationConstraints.py:263-392, StructureFactorConstraints.py:933-1096).
 * ---------------------------------------------------------------------------------- */
typedef struct frmc_store frmc_store;

/* Constraint-level constants for one model, prepared on the host with the reference's
 * own numpy expressions so the device epilogue can mirror them bit for bit.
 * All pointers are HOST pointers, copied at frmc_model_add time. */
typedef struct frmc_model_desc {
    int32_t kind;           /* FRMC_KIND_* */
    int32_t n_pairs;        /* element pairs in sorted(combinations_with_replacement) order */
    const int32_t *pair_a;  /* [n_pairs] element index idi (PairDistributionConstraints.py:864) */
    const int32_t *pair_b;  /* [n_pairs] element index idj */
    const float *pair_w;    /* [n_pairs] w_ij as float32 (:487-489) */
    const float *pair_D;    /* [n_pairs] D_ij = float32(N_ij / volume) (:867-874) */
    const float *shell_volumes; /* [hs]  (:758) */
    const float *prefactor;     /* [hs] (4.*PI*shellCenters*rho0) as numpy evaluates it (:881); unused for PCF unless scale!=1 */
    const float *shape;         /* [hs] or NULL, shape-function array subtracted (:883-884) */
    float scale;                /* fitted scale factor applied when != 1 (:886-888) */
    int32_t n_out;              /* length of the model total: hs (PDF/PCF) or nQ (SQ/RSQ) */
    const float *experimental;  /* [n_out] experimental data */
    const float *data_weights;  /* [n_out] or NULL (usedDataWeights, :835-838) */
    const float *gr2sq;         /* [hs,n_out] row-major Gr2SqMatrix (SQ/RSQ), else NULL (StructureFactorConstraints.py:302-312) */
    int32_t sq_exact;           /* 1: sequential fp32 accumulation over r (bit-exact vs numpy, :772-773); 0: split-r fast sum */
} frmc_model_desc;

frmc_store *frmc_store_create(int dev, int64_t n, const float *coords, const float *basis, int isPBC,
                              const int32_t *mol, const int32_t *el, int nEl);
void frmc_store_destroy(frmc_store *s);
/* re-upload all coordinates (set_pdb / set_boundary_conditions / external edits); invalidates data */
int frmc_store_set_coords(frmc_store *s, const float *coords, const float *basis);
int frmc_store_get_coords(frmc_store *s, float *coords_out);
/* CUDA stream the store launches on (a cudaStream_t), so callers can time with events on it */
void *frmc_store_stream(frmc_store *s);

/* register an r-grid (minDistance, maxDistance, bin, histSize); returns grid id >= 0 */
int frmc_grid_add(frmc_store *s, float rmin, float rmax, float bin, int hs);
/* register a model on a grid; returns model id >= 0 */
int frmc_model_add(frmc_store *s, int grid, const frmc_model_desc *desc);
int frmc_model_set_scale(frmc_store *s, int model, float scale);
/* replace (or, with NULL, drop) the shape-function array [histSize] an r-space model subtracts
 * (PairDistributionConstraints.py:882-883): what _update_shape_array does every shapeUpdateFreq accepted
 * moves (:316-343, :362-374).  Follow with frmc_finalize_data to refresh the committed chi^2. */
int frmc_model_set_shape(frmc_store *s, int model, const float *shape);
/* window function (set_window_function, PairDistributionConstraints.py:676-713; already normalised by the
 * caller as there): total = np.convolve(total, window, "same") after scale and prior (:892-893).  This branch
 * is within 1e-6 of numpy, not bit-exact (numpy's dot kernel fixes no summation order).  NULL / 0: off. */
int frmc_model_set_window(frmc_store *s, int model, const float *window, int n_window);
/* multiframe prior and weight (Core/Constraint.py:1160-1177): total = prior + weight * total, between the scale
 * factor and the window.  prior [n_out]; NULL: off. */
int frmc_model_set_multiframe_prior(frmc_store *s, int model, const float *prior, float weight);
/* Scale-factor refit (ExperimentalConstraint.set_adjust_scale_factor / fit_scale_factor /
 * get_adjusted_scale_factor, Core/Constraint.py:1363-1423): when frequency > 0 every evaluation made while
 * accepted % frequency == 0 fits SF = sum(w*M*E)/sum(M^2) (numpy fp32 pairwise order, on G(r) for the
 * r-space kinds and on S(Q)-1 for the Q-space kinds), clips it to [sf_min, sf_max] and scales the total
 * with it; frmc_accept makes the last used value the model's scale factor (accept_move,
 * PairDistributionConstraints.py:1150).  `accepted` is the engine's count of accepted moves: the store
 * counts its own frmc_accept calls, frmc_store_set_accepted re-bases it. */
int frmc_model_set_adjust(frmc_store *s, int model, int frequency, float sf_min, float sf_max);
int frmc_model_get_scale(frmc_store *s, int model, float *committed, float *last_used);
int frmc_store_set_accepted(frmc_store *s, uint64_t accepted);
/* Persistent per-move kernel (default off; FRMC_PERSISTENT=1 switches it on for new stores).  When on, frmc_propose /
 * frmc_step keep ONE cooperative kernel resident across a run of proposals and hand it each move through mapped
 * pinned memory, which removes the per-move kernel launch (about 8 us of the ~38 us round trip).  Results are
 * identical.  Every other entry point that touches the device ends the run first; the kernel also leaves by
 * itself after 2 ms without a proposal (FRMC_PERSIST_IDLE_US) and is restarted on demand.  Falls back to one
 * launch per proposal for models whose S(Q) slab does not fit shared memory or that refit their scale factor.
 * frmc_store_persistent_stats: kernels started and proposals served so far. */
int frmc_store_set_persistent(frmc_store *s, int on);
int frmc_store_persistent_stats(frmc_store *s, uint64_t *kernel_launches, uint64_t *commands);

/* compute_data: full histogram of every grid (tiled kernel), totals and chi^2 per model.
 * chi2 [n_models] fp32 (np.add.reduce result), may be NULL. */
int frmc_compute_data(frmc_store *s, float *chi2);
/* sharded variant for one-process-per-GPU runs: computes this caller's partial integer
 * histograms only (device resident); combine with frmc_grid_counts_ptr + an NCCL
 * allreduce issued by the host plumbing, then call frmc_finalize_data. */
int frmc_compute_data_shard(frmc_store *s, int shard, int nshards);
/* DEVICE pointer to grid's committed counts: int64 [2][nEl*nEl][hs] (0=intra,1=inter) */
void *frmc_grid_counts_ptr(frmc_store *s, int grid, int64_t *n_cells);
int frmc_finalize_data(frmc_store *s, float *chi2);

/* compute_before_move + compute_after_move fused: group `indexes` (original atom indices)
 * moves to `moved` ([k,3] box coordinates, not wrapped).  One pass over the store forms
 * after-minus-before for every grid, then totals and chi^2 per model.  The proposal stays
 * staged until frmc_accept / frmc_reject. */
int frmc_propose(frmc_store *s, const int32_t *indexes, int k, const float *moved, float *chi2_after);
int frmc_accept(frmc_store *s);
int frmc_reject(frmc_store *s);
/* One host call per Metropolis step: resolve the staged proposal (previous = 1 accept, 0 reject;
 * ignored when nothing is staged) and evaluate the next one. */
int frmc_step(frmc_store *s, int previous, const int32_t *indexes, int k, const float *moved, float *chi2_after);
/* ---- the distance-constraint pre-filter on the device store (the per-move pass shares the
 * store) -----------------------------------------------------------------------------------------------------------
 * InterMolecularDistanceConstraint / IntraMolecularDistanceConstraint evaluate, for the k atoms of a move,
 * M = multiple_atomic_distances_coords(indexes, all atoms) and F = full_atomic_distances_coords(the group alone) before
 * and after the move (Constraints/DistanceConstraints.py:606-737), and Engine.py:3281-3290 does so BEFORE the
 * experimental constraints on every step.  frmc_store_distance_add registers the constraint once on the store whose
 * atoms the histogram constraints move (type [n] = typesIndex by atom; lowerLimit / upperLimit [nT*nT] indexed
 * [type_i, type_a]; flags = FRMC_AD_*); frmc_store_distance_move evaluates all four quantities of one move in one pass
 * over the resident records -- no coordinate upload.  counts_out / sums_out: [4][2][nT*nT] = (M before, F before,
 * M after, F after) x (intra, inter) x [type_a, type_i]; the float32 sums are accumulated in the reference's loop
 * order (bit-identical).  frmc_store_move_atoms applies an accepted move on a store without histogram models (with
 * models, frmc_accept does). */
int frmc_store_distance_add(frmc_store *s, const int32_t *type, int nT, const float *lowerLimit, const float *upperLimit, int flags);
int frmc_store_distance_move(frmc_store *s, int id, const int32_t *indexes, int k, const float *moved, int32_t *counts_out,
                             float *sums_out);
int frmc_store_move_atoms(frmc_store *s, const int32_t *indexes, int k, const float *moved);

/* ---- the coordination-number pre-filter on the device store -----------------------------
 * AtomicCoordinationNumberConstraint.compute_before_move / compute_after_move
 * (Constraints/AtomicCoordinationConstraints.py:519-577) call multi_atoms_coord_number_coords
 * (Extensions/atomic_coordination.pyx:280-313 through :207-240) for the k atoms of a move, before and after it.
 * frmc_store_coordination_add registers the definitions once on the store whose atoms the histogram constraints move:
 * definition d has the core atoms core_indexes[core_offsets[d] .. core_offsets[d+1]) and the shell atoms
 * shell_indexes[shell_offsets[d] .. shell_offsets[d+1]) (atom indexes of the store's layout, each atom at most once per
 * list; at most 32 definitions) and the shell lower[d] <= distance <= upper[d] (both ends inclusive, the float32 distance
 * of pairs_distances_to_point; the atom itself is not skipped).  Returns the constraint's id (>= 0) or an error code.
 * frmc_store_coordination_move evaluates one move in ONE launch over the resident records -- no coordinate upload, no
 * list upload: counts_out [2][ndef] = what multi_atoms_coord_number_coords adds to coordNumData for `indexes` on the
 * stored coordinates (before) and on the coordinates with the group at `moved` (after).  Integer counts: exactly the
 * reference's float32 values while a cell stays below 2^24.  The whole-system count of compute_data stays with
 * frmc_coordination_counts. */
int frmc_store_coordination_add(frmc_store *s, int ndef, const int64_t *core_offsets, const int32_t *core_indexes,
                                const int64_t *shell_offsets, const int32_t *shell_indexes, const float *lower, const float *upper);
int frmc_store_coordination_move(frmc_store *s, int id, const int32_t *indexes, int k, const float *moved, int32_t *counts_out);

/* ---- dynamic N and persisted state -------------------------------------------------
 * Atom removal (Engine.__on_runtime_step_try_remove, Engine.py:3231-3276; compute_as_if_amputated / accept_amputation /
 * reject_amputation of the three constraints, PairDistributionConstraints.py:1168-1238,
 * PairCorrelationConstraints.py:394-462, StructureFactorConstraints.py:1098-1166; Engine._on_collector_collect_atom,
 * Engine.py:758-797) without rebuilding the store: the atom's row is subtracted from the running counts by one pass
 * over the store, the constraint-level constants of the system with one atom fewer are swapped in for the evaluation,
 * and on acceptance the record becomes padding.  Afterwards every `indexes` argument of this ABI is the engine's
 * RELATIVE index (its arrays are np.delete'd: atoms behind the removed one move down by one); frmc_store_get_coords
 * returns one row per remaining atom.
 *
 * frmc_amputation_desc: the constants one model uses while the atom is tried as removed, prepared by the host with the
 * reference's numpy expressions: pair_w [n_pairs] = the weighting scheme for numberOfAtomsPerElement[el] - 1
 * (:1190-1192), pair_D [n_pairs] = D_ij of those counts (:867-874), prefactor [histSize] = 4 pi r rho0 with
 * rho0 = (N - 1) / volume (:1198).  NULL members keep the model's own. */
typedef struct frmc_amputation_desc {
    const float *pair_w;
    const float *pair_D;
    const float *prefactor;
} frmc_amputation_desc;
/* compute_as_if_amputated for every model at once: chi2_out [n_models] = amputationStandardError.  descs: one per
 * model or NULL.  allow_fit = engine._RT_moveGenerator.allowFittingScaleFactor (0: no scale-factor refit in this
 * evaluation whatever the schedule says, :1195-1197).  The amputation stays staged until accepted or rejected. */
int frmc_propose_amputation(frmc_store *s, int32_t index, const frmc_amputation_desc *descs, int allow_fit, float *chi2_out);
/* accept_amputation + _on_collector_collect_atom: data = data - row, standardError = amputationStandardError, the
 * scale factor the evaluation used becomes the model's, the atom leaves the store (n drops by one).  The models keep
 * their OWN constants: follow with frmc_model_set_constants for what the engine's new state implies. */
int frmc_accept_amputation(frmc_store *s);
int frmc_reject_amputation(frmc_store *s);
/* replace a model's pair weights / D_ij / prefactor (NULL keeps one); the 3-op division is re-validated */
int frmc_model_set_constants(frmc_store *s, int model, const float *pair_w, const float *pair_D, const float *prefactor);
/* atoms the store holds now */
int64_t frmc_store_n_atoms(frmc_store *s);
/* Resume from persisted state (Constraint._dump_to_repository / the engine's runtime save, Core/Constraint.py:275-288,
 * Engine.py:1188-1202, store data["intra"] / data["inter"]): the saved float32 arrays [nEl,nEl,histSize] become the
 * grid's committed counts (every cell must hold an integer) instead of a full-histogram pass; follow with
 * frmc_finalize_data for totals and chi^2.  The inverse of frmc_export_data. */
int frmc_import_data(frmc_store *s, int grid, const float *hintra, const float *hinter);

/* A RUN of n proposals tried in sequence with the engine's own acceptance rule, resolved on the device
 * (replaces n rounds of Engine.__on_runtime_step_try_move, Engine.py:3302-3338, for the histogram constraints):
 *   total_new = sum_m chi2_m / variance_sq[m]           (Engine.compute_total_standard_error, Engine.py:3024-3029)
 *   total_new > total:  rejected when the NEXT pre-drawn random number > tolerance, else accepted ("tolerated")
 *   otherwise accepted; an accepted proposal makes total = total_new, moves its atoms and commits its histograms.
 * Proposal j moves group_sizes[j] atoms (NULL: one atom each); `indexes` / `moved` hold the groups back to back
 * (original atom indices, [k,3] box coordinates each).  rand: n numbers, consumed in order, one per worse proposal
 * (generate_random_float is only drawn for those); *n_rand_used returns how many were consumed.  *total_io: the engine's
 * totalStandardError before / after.  chi2_out [n][n_models] (the chi2 every proposal was judged on), decisions [n]
 * (0 rejected, 1 accepted, 2 accepted within the tolerance), n_rand_used and device_ms (CUDA-event time of the
 * launches) may be NULL.  The outcome is identical to n frmc_step calls with that rule on the host.  One pass over
 * the store serves up to 32 proposals / 64 moved atoms (16 B/atom of traffic for all of them), and ONE launch works
 * through all such batches of the run (up to 128 per launch); see batch_kernel in csrc/store.cu.  Scale-factor refit
 * schedules run inside the batch kernel; models with an S(Q) slab beyond shared memory run the same rule with one
 * launch per proposal. */
int frmc_run_batch(frmc_store *s, int n, const int32_t *group_sizes, const int32_t *indexes, const float *moved,
                   const float *variance_sq, float tolerance, const float *rand, float *total_io,
                   float *chi2_out, int32_t *decisions, int32_t *n_rand_used, double *device_ms);
/* ---- device-generated runs of moves (move generation, transform_coordinates and move
 * application on the device, so that a run of steps needs no per-step host<->device traffic at all) --------------
 * What Engine.run does per step for groups moved by a TranslationGenerator under a RandomSelector -- select a group
 * (Engine.py:3168), translate its atoms' real coordinates by a random vector (Generators/Translations.py:171-187,
 * Core/Collection.py:674-701), transform_coordinates to box coordinates (Engine.py:3222-3223), evaluate, decide
 * (Engine.py:3302-3338), and on acceptance update realCoordinates and boxCoordinates (:3337-3338) -- for a whole RUN of
 * steps on the device.  The random numbers are COUNTER BASED (fullrmc_b200/rng.py states the contract: Philox4x32-10,
 * key = seed, counter = step number; group index, direction, amplitude and the acceptance number of step c are pure
 * functions of (seed, c)): the reference's own Mersenne-Twister streams are consumed in an order that depends on every
 * earlier decision, so they cannot be drawn ahead.  A reference Engine equipped with the selector / generator plug-ins
 * of fullrmc_b200/engine_plugins.py draws the same numbers and walks the same trajectory, bit for bit
 * (tests/test_generated_runs.py).
 *   frmc_store_set_real_coords  engine.realCoordinates [n,3] and engine.reciprocalBasisVectors [3,3] (both NULL for a
 *                               non-periodic store, whose box coordinates are the real ones)
 *   frmc_store_set_groups       the engine's groups: atoms indexes[offsets[g] .. offsets[g+1]) of group g
 *   frmc_run_generated          n steps numbered first_counter ..; amplitude range [amp_min, amp_max) as
 *                               TranslationGenerator.amplitude; the other arguments as frmc_run_batch.  groups_out [n]
 *                               and rand_out [n] (the acceptance numbers) may be NULL.
 * Moves accepted through any other entry point invalidate the real coordinates (set them again). */
int frmc_store_set_real_coords(frmc_store *s, const float *real, const float *rbasis);
int frmc_store_get_real_coords(frmc_store *s, float *real_out);
int frmc_store_set_groups(frmc_store *s, int n_groups, const int32_t *offsets, const int32_t *indexes);
int frmc_run_generated(frmc_store *s, int n, uint64_t seed, uint64_t first_counter, float amp_min, float amp_max,
                       const float *variance_sq, float tolerance, float *total_io, float *chi2_out, int32_t *decisions,
                       int32_t *groups_out, float *rand_out, double *device_ms);
/* batch launches, evaluation rounds inside them and proposals they resolved so far */
int frmc_store_batch_stats(frmc_store *s, uint64_t *launches, uint64_t *rounds, uint64_t *proposals);
/* Debug (FRMC_BATCH_STAMPS=1 in the environment before the first run): globaltimer ns of CTA 0 at the phase
 * boundaries of the LAST batch launch: [0] start, [1] buffers cleared, [2] delta pass done, [3] end, then per round r
 * [4+5r..]: own epilogue done, all epilogues done, decisions made, own commit share done, commit visible (0: no commit). */
int frmc_store_batch_stamps(frmc_store *s, int64_t *out, int n);
/* chi2 per model of the committed state (constraint.standardError) */
int frmc_store_committed_chi2(frmc_store *s, float *chi2);
/* Measurement helper: re-launch the staged proposal's device pipeline `reps` times back to back
 * (no host round trip in between) and return the average device time per launch, CUDA events on
 * the store's stream.  The staged state is restored afterwards. */
int frmc_store_replay_proposal(frmc_store *s, int reps, double *ms_per_launch);

/* export committed histograms as the reference's data["intra"], data["inter"] (fp32 [nEl,nEl,hs]) */
int frmc_export_data(frmc_store *s, int grid, float *hintra, float *hinter);
/* export the staged (after-move) or committed model total (length n_out) */
int frmc_export_total(frmc_store *s, int model, int staged, float *out);
/* events where the fp32 bin index rounded up to histSize (dropped), summed over the store's life */
uint64_t frmc_store_edge_overflow(frmc_store *s);
/* distance evaluations (pairs of records actually swept, padding included) of the last
 * frmc_compute_data / _shard on this store; n(n-1)/2 is the reference's count */
uint64_t frmc_store_swept_pairs(frmc_store *s);
/* Optional per-kernel timing with CUDA events on the store's stream (bench.py's roofline leg).
 * which: 0 = per-move delta pass, 1 = full-histogram kernel, 2 = epilogue (G(r)/S(Q)/chi^2 kernels),
 * 3 = commit/clear kernels.  get_timing synchronises the stream and returns the accumulated
 * device time in ms and the number of timed launches since timing was switched on. */
int frmc_store_set_timing(frmc_store *s, int on);
int frmc_store_get_timing(frmc_store *s, int which, double *ms_total, uint64_t *launches);
/* Debug: clock64() stamps of the last epilogue launch, [model][8]: 0 start, 1 pair table staged,
 * 2 r-space function done, 3 S(Q) slice done, 4 ticket taken, 5 chi^2 done (CTA x=0 of each model). */
int frmc_store_debug_stamps(frmc_store *s, int64_t *out, int n);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t frmc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FULLRMC_B200_H */
