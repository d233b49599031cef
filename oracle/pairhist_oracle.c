/*
 * oracle/pairhist_oracle.c -- CPU restatement of fullrmc's pair-histogram hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (fullrmc_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the
 * reference's own compiled Cython modules (oracle/_ref, built by oracle/build_ref.py
 * from the .pyx files of /root/reference/Extensions) in tests/test_oracle.py, and against the
 * golden vectors under tests/golden/ that those modules generated.
 *
 * Each function cites the reference lines it restates.  Arithmetic is strict fp32
 * evaluated left to right with no FMA contraction (build with -ffp-contract=off),
 * exactly what gcc -O2 produces from the Cython-generated C on x86-64/SSE2.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* pairs_distances.pyx:31-32 -- round half away from zero, NOT rintf:
 *   floor(num + 0.5) if num > 0 else ceil(num - 0.5)
 * the float add happens in fp32 before the (exact) double floor/ceil. */
static inline float orc_round(float num)
{
    return (num > 0.0f) ? floorf(num + 0.5f) : ceilf(num - 0.5f);
}

/* pairs_distances.pyx:348-386 (_boxcoords_realdistances_to_indexcoords_PBC) and
 * :326-341 (to_boxpoint variant): distance from fractional point p to fractional
 * coords row c under the lattice `basis` (rows = lattice vectors). */
static inline float orc_dist_pbc(float px, float py, float pz, const float *c, const float *b)
{
    float diff_x = px - c[0];
    float diff_y = py - c[1];
    float diff_z = pz - c[2];
    float box_dx = diff_x - orc_round(diff_x);
    float box_dy = diff_y - orc_round(diff_y);
    float box_dz = diff_z - orc_round(diff_z);
    float real_dx = box_dx * b[0] + box_dy * b[3] + box_dz * b[6];
    float real_dy = box_dx * b[1] + box_dy * b[4] + box_dz * b[7];
    float real_dz = box_dx * b[2] + box_dy * b[5] + box_dz * b[8];
    return sqrtf(real_dx * real_dx + real_dy * real_dy + real_dz * real_dz);
}

/* pairs_distances.pyx:443-471 (_realcoords_realdistances_to_indexcoords_IBC) */
static inline float orc_dist_ibc(float px, float py, float pz, const float *c)
{
    float real_dx = px - c[0];
    float real_dy = py - c[1];
    float real_dz = pz - c[2];
    return sqrtf(real_dx * real_dx + real_dy * real_dy + real_dz * real_dz);
}

/* pairs_distances.pyx:874-916 (pairs_distances_to_indexcoords).  Entries below the
 * start index are left untouched (the reference leaves them uninitialised). */
void orc_pairs_distances_to_indexcoords(int32_t atomIndex, const float *coords, int64_t n,
                                        const float *basis, int isPBC, int allAtoms, float *distances)
{
    const float px = coords[3 * (int64_t)atomIndex + 0];
    const float py = coords[3 * (int64_t)atomIndex + 1];
    const float pz = coords[3 * (int64_t)atomIndex + 2];
    int64_t start = allAtoms ? 0 : atomIndex;
    for (int64_t i = start; i < n; ++i)
        distances[i] = isPBC ? orc_dist_pbc(px, py, pz, coords + 3 * i, basis)
                             : orc_dist_ibc(px, py, pz, coords + 3 * i);
}

/* pairs_distances.pyx:827-866 (pairs_distances_to_point): IBC variant computes
 * coords - point (sign irrelevant to the distance). */
void orc_pairs_distances_to_point(const float *point, const float *coords, int64_t n,
                                  const float *basis, int isPBC, float *distances)
{
    for (int64_t i = 0; i < n; ++i)
        distances[i] = isPBC ? orc_dist_pbc(point[0], point[1], point[2], coords + 3 * i, basis)
                             : orc_dist_ibc(point[0], point[1], point[2], coords + 3 * i);
}

/* pairs_distances.pyx:580-617 / :626-666 (pairs_differences_to_point / _to_indexcoords).
 * PBC: point - coords[i] wrapped then multiplied by basis (:142-165, :201-235);
 * IBC: point - coords[i] (:173-192, :244-270).  ibcPointMinusCoords=0 gives the opposite
 * sign (the IBC *distance* kernel :414-434 subtracts that way; irrelevant to a norm). */
void orc_pairs_differences(const float *point, const float *coords, int64_t n, const float *b,
                           int isPBC, int ibcPointMinusCoords, int64_t start, float *diffs)
{
    for (int64_t i = start; i < n; ++i) {
        const float *c = coords + 3 * i;
        if (isPBC) {
            float diff_x = point[0] - c[0], diff_y = point[1] - c[1], diff_z = point[2] - c[2];
            float box_dx = diff_x - orc_round(diff_x);
            float box_dy = diff_y - orc_round(diff_y);
            float box_dz = diff_z - orc_round(diff_z);
            diffs[3 * i + 0] = box_dx * b[0] + box_dy * b[3] + box_dz * b[6];
            diffs[3 * i + 1] = box_dx * b[1] + box_dy * b[4] + box_dz * b[7];
            diffs[3 * i + 2] = box_dx * b[2] + box_dy * b[5] + box_dz * b[8];
        } else if (ibcPointMinusCoords) {
            diffs[3 * i + 0] = point[0] - c[0];
            diffs[3 * i + 1] = point[1] - c[1];
            diffs[3 * i + 2] = point[2] - c[2];
        } else {
            diffs[3 * i + 0] = c[0] - point[0];
            diffs[3 * i + 1] = c[1] - point[1];
            diffs[3 * i + 2] = c[2] - point[2];
        }
    }
}

/* pairs_histograms.pyx:36-68 (_single_pairs_histograms): the bin rule.
 *   skip j == atom; skip d < min; skip d >= max; bin = (int)((d - min) / bin)
 * The reference does no bounds check (boundscheck(False)); a bin index that lands
 * on histSize because of fp32 rounding is undefined behaviour there.  Here it is
 * dropped and counted in *overflow so the GPU path can report the same events. */
/* orc_emulate_spill != 0: reproduce what the reference's unchecked write actually does on this
 * platform when bin == histSize: the flat index (a*nEl+b)*hs + bin lands in the NEXT slab's first
 * bins (same array); events that would leave the array are dropped.  Either way the event is
 * counted in *overflow.  Default 0: drop (the physically meaningful behaviour). */
static int orc_emulate_spill = 0;
void orc_set_emulate_spill(int on) { orc_emulate_spill = on; }

static inline void orc_bin_one(float d, int same_mol, int32_t ea, int32_t eb, int nEl, int hs,
                               float rmin, float rmax, float bin, float *hintra, float *hinter,
                               uint64_t *overflow)
{
    if (d < rmin) return;
    if (d >= rmax) return;
    int32_t b = (int32_t)((d - rmin) / bin);
    int64_t at = ((int64_t)ea * nEl + eb) * hs + b;
    if (b >= hs || b < 0) {
        if (overflow) (*overflow)++;
        if (!orc_emulate_spill || b < 0 || at >= (int64_t)nEl * nEl * hs) return;
    }
    if (same_mol) hintra[at] += 1.0f; else hinter[at] += 1.0f;
}

/* pairs_histograms.pyx:77-141 (single_pairs_histograms): in-place update from a
 * precomputed distance row with stride `dstride` (distances[:,i] views, :270). */
void orc_single_pairs_histograms(int32_t atomIndex, const float *distances, int64_t dstride, int64_t n,
                                 const int32_t *mol, const int32_t *el, int nEl, int hs,
                                 float *hintra, float *hinter, float rmin, float rmax, float bin,
                                 int allAtoms, uint64_t *overflow)
{
    const int32_t am = mol[atomIndex], ae = el[atomIndex];
    int64_t start = allAtoms ? 0 : atomIndex;
    for (int64_t i = start; i < n; ++i) {
        if (i == atomIndex) continue;
        orc_bin_one(distances[i * dstride], mol[i] == am, ae, el[i], nEl, hs, rmin, rmax, bin,
                    hintra, hinter, overflow);
    }
}

static void orc_row(int32_t a, const float *coords, int64_t n, const float *basis, int isPBC,
                    const int32_t *mol, const int32_t *el, int nEl, int hs, float rmin, float rmax,
                    float bin, int allAtoms, float *hintra, float *hinter, uint64_t *overflow)
{
    const float px = coords[3 * (int64_t)a], py = coords[3 * (int64_t)a + 1], pz = coords[3 * (int64_t)a + 2];
    const int32_t am = mol[a], ae = el[a];
    int64_t start = allAtoms ? 0 : a;
    for (int64_t i = start; i < n; ++i) {
        if (i == a) continue;
        float d = isPBC ? orc_dist_pbc(px, py, pz, coords + 3 * i, basis)
                        : orc_dist_ibc(px, py, pz, coords + 3 * i);
        orc_bin_one(d, mol[i] == am, ae, el[i], nEl, hs, rmin, rmax, bin, hintra, hinter, overflow);
    }
}

/* pairs_histograms.pyx:150-217 (multiple_pairs_histograms_coords): for each listed atom,
 * one distance row (pairs_distances_to_indexcoords) + one single_pairs_histograms.
 * hintra/hinter must be zeroed by the caller (the reference allocates np.zeros).
 * nthreads > 1 splits the listed atoms over OpenMP threads with thread-private
 * histograms summed at the end (exact while every cell < 2^24, like the fp32
 * increments of the reference itself). */
void orc_multiple_pairs_histograms_coords(const int32_t *indexes, int64_t k, const float *coords, int64_t n,
                                          const float *basis, int isPBC, const int32_t *mol,
                                          const int32_t *el, int nEl, float rmin, float rmax, float bin,
                                          int hs, int allAtoms, float *hintra, float *hinter,
                                          uint64_t *overflow, int nthreads)
{
    const int64_t cells = (int64_t)nEl * nEl * hs;
    uint64_t ov_total = 0;
    if (nthreads <= 1) {
        for (int64_t t = 0; t < k; ++t)
            orc_row(indexes[t], coords, n, basis, isPBC, mol, el, nEl, hs, rmin, rmax, bin, allAtoms,
                    hintra, hinter, &ov_total);
    } else {
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads) reduction(+ : ov_total)
#endif
        {
            float *pi = (float *)calloc((size_t)cells, sizeof(float));
            float *pe = (float *)calloc((size_t)cells, sizeof(float));
            uint64_t ov = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
            for (int64_t t = 0; t < k; ++t)
                orc_row(indexes[t], coords, n, basis, isPBC, mol, el, nEl, hs, rmin, rmax, bin, allAtoms,
                        pi, pe, &ov);
#ifdef _OPENMP
#pragma omp critical
#endif
            {
                for (int64_t c = 0; c < cells; ++c) { hintra[c] += pi[c]; hinter[c] += pe[c]; }
            }
            ov_total += ov;
            free(pi); free(pe);
        }
    }
    if (overflow) *overflow += ov_total;
}

/* pairs_histograms.pyx:289-335 (full_pairs_histograms_coords): indexes = arange(N),
 * allAtoms=False, i.e. the ordered upper triangle [el[i], el[j]] with i < j. */
void orc_full_pairs_histograms_coords(const float *coords, int64_t n, const float *basis, int isPBC,
                                      const int32_t *mol, const int32_t *el, int nEl, float rmin,
                                      float rmax, float bin, int hs, float *hintra, float *hinter,
                                      uint64_t *overflow, int nthreads)
{
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) idx[i] = (int32_t)i;
    orc_multiple_pairs_histograms_coords(idx, n, coords, n, basis, isPBC, mol, el, nEl, rmin, rmax, bin,
                                         hs, 0, hintra, hinter, overflow, nthreads);
    free(idx);
}

/* pairs_histograms.pyx:225-281 / :343-383 (multiple/full_pairs_histograms_dists):
 * distances is [N, k] row-major; column t belongs to indexes[t]. */
void orc_multiple_pairs_histograms_dists(const int32_t *indexes, int64_t k, const float *distances,
                                         int64_t n, const int32_t *mol, const int32_t *el, int nEl,
                                         float rmin, float rmax, float bin, int hs, int allAtoms,
                                         float *hintra, float *hinter, uint64_t *overflow)
{
    for (int64_t t = 0; t < k; ++t)
        orc_single_pairs_histograms(indexes[t], distances + t, k, n, mol, el, nEl, hs, hintra, hinter,
                                    rmin, rmax, bin, allAtoms, overflow);
}

/* reciprocal_space.pyx:82-109 (Gr_to_sq): every term is evaluated in double through
 * Python-level np.sin (the operands q*r and dr are fp32 products/differences promoted
 * to double), ROUNDED TO FP32 (the generated C converts the Python float with
 * __Pyx_PyFloat_AsFloat before the +=), then added in fp32 to sq[qidx] which starts at 1. */
void orc_Gr_to_sq(const float *distances, const float *Gr, int64_t n, const float *qrange, int64_t m, float *sq)
{
    float dr = distances[1] - distances[0];
    for (int64_t qi = 0; qi < m; ++qi) {
        float q = qrange[qi];
        float acc = 1.0f;
        for (int64_t ri = 0; ri < n; ++ri) {
            float r = distances[ri];
            double term = (double)dr * (sin((double)(q * r)) / (double)q) * (double)Gr[ri];
            acc += (float)term;
        }
        sq[qi] = acc;
    }
}

/* reciprocal_space.pyx:42-73 (gr_to_sq):
 *   sq[q] += fact * ( dr*r*(np.sin(q*r)/q)*(gr[r]-1) ),  fact = 4*pi32*rho in fp32,
 * dr*r is an fp32 product; gr-1.0, the sine term and the products are double; the term
 * is rounded to fp32 before the fp32 +=. */
void orc_gr_to_sq(const float *distances, const float *gr, int64_t n, const float *qrange, int64_t m,
                  float rho, float *sq)
{
    float dr = distances[1] - distances[0];
    float fact = 4.0f * 3.1415927f * rho;
    for (int64_t qi = 0; qi < m; ++qi) {
        float q = qrange[qi];
        float acc = 1.0f;
        for (int64_t ri = 0; ri < n; ++ri) {
            float r = distances[ri];
            double term = (double)fact * (((double)(dr * r) * (sin((double)(q * r)) / (double)q)) * ((double)gr[ri] - 1.0));
            acc += (float)term;
        }
        sq[qi] = acc;
    }
}

/* ---------------------------------------------------------------- atomic distances
 * Extensions/atomic_distances.pyx:42-118 (_single_atomic_distances_dists) driven by
 * multiple_atomic_distances_coords (:326-417) / full_atomic_distances_coords (:500-567): for every listed atom a
 * and every other atom i >= start, a pair whose distance falls inside (countWithinLimits) or outside the
 * [lower, upper) window of its type pair adds its (optionally reduced) distance to dintra/dinter[type_a, type_i]
 * and one to nintra/ninter, in exactly this loop order (the float sums depend on it).  Limits are indexed
 * [type_i, type_a], outputs [type_a, type_i].  flags: bit0 interMolecular, bit1 intraMolecular,
 * bit2 countWithinLimits, bit3 reduceDistanceToUpper, bit4 reduceDistanceToLower, bit5 reduceDistance. */
void orc_multiple_atomic_distances_coords(const int32_t *indexes, int64_t k, const float *coords, int64_t n,
                                          const float *basis, int isPBC, const int32_t *mol, const int32_t *el, int nT,
                                          const float *lowerLimit, const float *upperLimit, int flags, int allAtoms,
                                          int32_t *nintra, float *dintra, int32_t *ninter, float *dinter)
{
    const int inter = flags & 1, intra = (flags >> 1) & 1, within = (flags >> 2) & 1;
    const int toUpper = (flags >> 3) & 1, toLower = (flags >> 4) & 1, reduce = (flags >> 5) & 1;
    for (int64_t t = 0; t < k; ++t) {
        const int32_t a = indexes[t];
        const float px = coords[3 * (int64_t)a], py = coords[3 * (int64_t)a + 1], pz = coords[3 * (int64_t)a + 2];
        const int32_t am = mol[a], ae = el[a];
        const int64_t start = allAtoms ? 0 : a;
        for (int64_t i = start; i < n; ++i) {
            if (i == a) continue;
            const int32_t im = mol[i];
            if (!intra && im == am) continue;
            if (!inter && im != am) continue;
            float distance = isPBC ? orc_dist_pbc(px, py, pz, coords + 3 * i, basis) : orc_dist_ibc(px, py, pz, coords + 3 * i);
            const int32_t ie = el[i];
            const float lower = lowerLimit[ie * nT + ae], upper = upperLimit[ie * nT + ae];
            if (within) {
                if (distance < lower) continue;
                if (distance >= upper) continue;
            } else if (distance >= lower && distance < upper) {
                continue;
            }
            if (toUpper) distance = (float)fabs(upper - distance);
            else if (toLower) distance = (float)fabs(lower - distance);
            else if (reduce) {
                if (distance > (lower + upper) / 2.0f) distance = (float)fabs(upper - distance);
                else distance = (float)fabs(lower - distance);
            }
            if (im == am) { dintra[ae * nT + ie] += distance; nintra[ae * nT + ie] += 1; }
            else { dinter[ae * nT + ie] += distance; ninter[ae * nT + ie] += 1; }
        }
    }
}

/* ---------------------------------------------------------------- coordination numbers
 * Extensions/atomic_coordination.pyx:89-108 (single_atom_single_shell_coords): distances from the core atom to
 * boxCoords[shellIndexes] by pairs_distances_to_point, then the counting loop :31-48 -- a float32 counter that
 * gains 1.0 for every distance with lower <= d <= upper (both ends inclusive; the core atom itself is not
 * skipped when it is in the list).  Negative indexes wrap as numpy fancy indexing does. */
float orc_single_atom_single_shell_coords(int32_t coreIndex, const int32_t *shellIndexes, int64_t ns, const float *coords,
                                          int64_t n, const float *basis, int isPBC, float lowerShell, float upperShell)
{
    const int64_t a = coreIndex < 0 ? coreIndex + n : coreIndex;
    const float px = coords[3 * a], py = coords[3 * a + 1], pz = coords[3 * a + 2];
    float coordNumber = 0.0f;
    for (int64_t i = 0; i < ns; ++i) {
        const int64_t j = shellIndexes[i] < 0 ? shellIndexes[i] + n : shellIndexes[i];
        const float d = isPBC ? orc_dist_pbc(px, py, pz, coords + 3 * j, basis) : orc_dist_ibc(px, py, pz, coords + 3 * j);
        if (lowerShell <= d && d <= upperShell) coordNumber += 1.0f;
    }
    return coordNumber;
}

/* atomic_coordination.pyx:71-85 (single_atom_single_shell_totdists; shellIndexes NULL = :55-67 subdists) */
float orc_single_atom_single_shell_dists(const float *distances, int64_t n, const int32_t *shellIndexes, int64_t ns,
                                         float lowerShell, float upperShell)
{
    float coordNumber = 0.0f;
    for (int64_t i = 0; i < ns; ++i) {
        const int64_t j = shellIndexes ? (shellIndexes[i] < 0 ? shellIndexes[i] + n : shellIndexes[i]) : i;
        const float d = distances[j];
        if (lowerShell <= d && d <= upperShell) coordNumber += 1.0f;
    }
    return coordNumber;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
