// stateless.cu -- signature-parity entry points for the reference's extension functions
// (Extensions/pairs_distances.pyx, pairs_histograms.pyx, reciprocal_space.pyx).
// Host buffers in, host buffers out; the kernels run on the per-device context stream.
#include "common.cuh"
#include "rng.cuh"
#include "layout.h"

#include <cstring>
#include <vector>

namespace frmc {

// ------------------------------------------------------------------ distances / differences
// One thread per (coords row i, point t).  Restates the cdef kernels of
// pairs_distances.pyx:201-471 (difference and distance, PBC and IBC, to-point and
// to-index variants).
template <bool PBC, bool WANT_DIFF>
__global__ void points_to_coords_kernel(const float *__restrict__ points, const long long *__restrict__ start,
                                        int k, const float *__restrict__ coords, long long n, Lattice L,
                                        int ibc_sign, float *__restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int t = blockIdx.y;
    if (i >= n || t >= k) return;
    long long s = start ? start[t] : 0;
    float rx = 0.f, ry = 0.f, rz = 0.f;
    bool live = i >= s;
    if (live) {
        float px = points[3 * t], py = points[3 * t + 1], pz = points[3 * t + 2];
        float cx = coords[3 * i], cy = coords[3 * i + 1], cz = coords[3 * i + 2];
        if (!PBC && ibc_sign < 0) {
            // coords - point (pairs_distances.pyx:264-272, :428-434)
            rx = __fsub_rn(cx, px); ry = __fsub_rn(cy, py); rz = __fsub_rn(cz, pz);
        } else {
            diff3<PBC>(px, py, pz, cx, cy, cz, L, rx, ry, rz);
        }
    }
    if (WANT_DIFF) {
        out[(i * 3 + 0) * k + t] = rx;
        out[(i * 3 + 1) * k + t] = ry;
        out[(i * 3 + 2) * k + t] = rz;
    } else {
        float d2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
        out[i * k + t] = live ? __fsqrt_rn(d2) : 0.f;
    }
}

// pairs_distances.pyx:73-96 / :119-134 (_from_to_*_realdifferences): to[i]-from[i]
template <bool PBC>
__global__ void from_to_kernel(const float *__restrict__ from, const float *__restrict__ to, long long n,
                               Lattice L, float *__restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float rx, ry, rz;
    diff3<PBC>(to[3 * i], to[3 * i + 1], to[3 * i + 2], from[3 * i], from[3 * i + 1], from[3 * i + 2], L, rx, ry, rz);
    out[3 * i] = rx; out[3 * i + 1] = ry; out[3 * i + 2] = rz;
}

// ------------------------------------------------------------------ rows histogram
// multiple_pairs_histograms_coords (pairs_histograms.pyx:150-217): each listed atom a
// against coords rows [start_a, n), j != a, ordered slab [el[a], el[j]].
// Rows are staged in shared memory (chunks of ROWS_PER_CHUNK on blockIdx.y), coords are
// streamed once per chunk with coalesced loads; counts go to global u32 cells.
struct RowRec { float x, y, z; int mol; int el; int index; int start; int pad; };
static const int ROWS_PER_CHUNK = 128;

template <int MODE>
__global__ void __launch_bounds__(256)
rows_hist_kernel(const float *__restrict__ coords, const int *__restrict__ mol, const int *__restrict__ el,
                 long long n, const int *__restrict__ idx, int k, int allAtoms, Lattice L, GridParams g, int nEl,
                 unsigned int *__restrict__ counts, unsigned long long *__restrict__ overflow)
{
    __shared__ RowRec rows[ROWS_PER_CHUNK];
    int r0 = blockIdx.y * ROWS_PER_CHUNK;
    int nr = min(ROWS_PER_CHUNK, k - r0);
    for (int t = threadIdx.x; t < nr; t += blockDim.x) {
        int a = idx[r0 + t];
        RowRec r;
        r.x = coords[3 * (long long)a]; r.y = coords[3 * (long long)a + 1]; r.z = coords[3 * (long long)a + 2];
        r.mol = mol[a]; r.el = el[a]; r.index = a; r.start = allAtoms ? 0 : a; r.pad = 0;
        rows[t] = r;
    }
    __syncthreads();
    const long long cells = (long long)nEl * nEl * g.hs;
    unsigned long long ov = 0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        float xj = coords[3 * j], yj = coords[3 * j + 1], zj = coords[3 * j + 2];
        int mj = mol[j], ej = el[j];
        for (int t = 0; t < nr; ++t) {
            const RowRec &r = rows[t];
            if (j < r.start || j == r.index) continue;
            float d2 = dist2<MODE>(r.x, r.y, r.z, xj, yj, zj, L);
            if (in_range(d2, g)) {
                int b = bin_index(d2, g);
                long long flat = ((long long)r.el * nEl + ej) * g.hs + b;
                if (b >= g.hs) ++ov;
                if (b < g.hs || (g.spill && flat < cells)) atomicAdd(&counts[flat + ((mj == r.mol) ? 0 : cells)], 1u);
            }
        }
    }
    if (ov) atomicAdd(overflow, ov);
}

// multiple_pairs_histograms_dists / single_pairs_histograms (pairs_histograms.pyx:36-68,
// :225-281): bin rule on precomputed distances, element (i,t) at distances[i*dstride_row + t*dstride_col].
__global__ void dists_hist_kernel(const float *__restrict__ distances, long long n, long long stride_row,
                                  long long stride_col, const int *__restrict__ mol, const int *__restrict__ el,
                                  const int *__restrict__ idx, int k, int allAtoms, GridParams g, int nEl,
                                  unsigned int *__restrict__ counts, unsigned long long *__restrict__ overflow)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int t = blockIdx.y;
    if (i >= n || t >= k) return;
    int a = idx[t];
    if (i == a) return;
    if (!allAtoms && i < a) return;
    float d = distances[i * stride_row + t * stride_col];
    int b;
    if (!bin_of_distance(d, g, b)) return;
    const long long cells = (long long)nEl * nEl * g.hs;
    const long long flat = ((long long)el[a] * nEl + el[i]) * g.hs + b;
    if (b >= g.hs || b < 0) {
        atomicAdd(overflow, 1ull);
        if (!g.spill || b < 0 || flat >= cells) return;
    }
    atomicAdd(&counts[flat + ((mol[i] == mol[a]) ? 0 : cells)], 1u);
}

// counts (u32) -> fp32, optionally added onto an existing fp32 histogram (in-place API)
__global__ void counts_to_float_kernel(const unsigned int *__restrict__ counts, const float *__restrict__ base,
                                       float *__restrict__ out, long long cells)
{
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    float v = (float)counts[c];
    out[c] = base ? __fadd_rn(base[c], v) : v;
}

// ------------------------------------------------------------------ reciprocal space
// reciprocal_space.pyx:82-109 / :42-73.  One thread per Q value, r in index order; each
// term is formed in double, rounded to fp32, then added to the fp32 accumulator
// (that is what the generated C does: __Pyx_PyFloat_AsFloat before the +=).
template <bool SMALL_G>
__global__ void r_to_q_kernel(const float *__restrict__ r, const float *__restrict__ y, long long n,
                              const float *__restrict__ q, long long m, float fact, float *__restrict__ sq)
{
    long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= m) return;
    float dr = __fsub_rn(r[1], r[0]);
    float qq = q[qi];
    float acc = 1.0f;
    for (long long ri = 0; ri < n; ++ri) {
        float rr = r[ri];
        double s = sin((double)__fmul_rn(qq, rr)) / (double)qq;
        double term;
        if (SMALL_G)   // gr_to_sq: fact * ( dr*r*(sin(qr)/q)*(gr-1) )
            term = (double)fact * (((double)__fmul_rn(dr, rr) * s) * ((double)y[ri] - 1.0));
        else           // Gr_to_sq: dr*(sin(qr)/q)*Gr
            term = ((double)dr * s) * (double)y[ri];
        acc = __fadd_rn(acc, (float)term);
    }
    sq[qi] = acc;
}

// reciprocal_space.pyx:118-145 sq_to_Gr, documented math (the reference body raises):
//   G(r) = (2/pi) * sum_q  q*(S(q)-1) * dq*sin(q*r)       fp32 operands, double sine, fp32 sum
__global__ void q_to_r_kernel(const float *__restrict__ q, const float *__restrict__ r, const float *__restrict__ sq,
                              long long m, long long n, float *__restrict__ Gr)
{
    long long ri = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= n) return;
    float dq = __fsub_rn(q[1], q[0]);
    float rr = r[ri];
    float acc = 0.0f;
    for (long long qi = 0; qi < m; ++qi) {
        float qq = q[qi];
        float qsq1 = __fmul_rn(qq, __fsub_rn(sq[qi], 1.0f));
        float sdq = __fmul_rn(dq, (float)sin((double)__fmul_rn(qq, rr)));
        acc = __fadd_rn(acc, __fmul_rn(qsq1, sdq));
    }
    Gr[ri] = (float)((2.0 / 3.14159265358979323846) * (double)acc);
}

// Shape function of a finite system (ShapeFunction, Constraints/Collection.py:20-125; refreshed by
// PairDistributionConstraint._update_shape_array, PairDistributionConstraints.py:316-343): from the ordered pair
// histograms of the whole system on the shape function's own coarse r-grid to G_shape(r) on the constraint's r values,
//     g(b)   = sum_pairs w_ij * (n_ij(b) / V_shell(b)) / D_ij          D_ij = N_ij / volume
//     G(b)   = 4 pi r_b rho0 (g(b) - 1)
//     S(q)-1 = sum_b G(b) * dr * sin(q r_b) / q                        (StructureFactorConstraints.py:302-312, 772-773)
//     G_s(r) = (2/pi) * dq * sum_q q (S(q)-1) sin(q r)                 (Collection.py:83-91)
// three small kernels, every sum in double (the reference's numpy float32 sums are within 1e-6 of it norm-wise;
// tests/test_golden_constraints.py holds the SiOx refreshes to that bar).
__global__ void shape_gr_kernel(const float *__restrict__ hintra, const float *__restrict__ hinter, int nEl, int hs, int n_pairs,
                                const int *__restrict__ pair_a, const int *__restrict__ pair_b, const double *__restrict__ pair_coef,
                                const float *__restrict__ shell_volumes, const float *__restrict__ shell_centers, double rho0,
                                double *__restrict__ G)
{
    const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (b >= hs) return;
    double g = 0.0;
    for (int p = 0; p < n_pairs; ++p) {
        const int ea = pair_a[p], eb = pair_b[p];
        const size_t ab = ((size_t)ea * nEl + eb) * hs + b, ba = ((size_t)eb * nEl + ea) * hs + b;
        double nij = (double)hintra[ab] + (double)hinter[ab];
        if (ea != eb) nij += (double)hintra[ba] + (double)hinter[ba];
        g += pair_coef[p] * nij / (double)shell_volumes[b];
    }
    G[b] = 4.0 * 3.14159265358979323846 * (double)shell_centers[b] * rho0 * (g - 1.0);
}

__global__ void shape_sq_kernel(const double *__restrict__ G, const float *__restrict__ shell_centers, int hs,
                                const float *__restrict__ q, int nq, double *__restrict__ S1)
{
    const int m = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (m >= nq) return;
    const double dr = (double)shell_centers[1] - (double)shell_centers[0], qq = (double)q[m];
    double acc = 0.0;
    for (int b = 0; b < hs; ++b) acc += G[b] * dr * sin(qq * (double)shell_centers[b]) / qq;
    S1[m] = acc;
}

__global__ void shape_back_kernel(const double *__restrict__ S1, const float *__restrict__ q, int nq, const float *__restrict__ r,
                                  int nr, float *__restrict__ out)
{
    const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k >= nr) return;
    const double dq = (double)q[1] - (double)q[0], rr = (double)r[k];
    double acc = 0.0;
    for (int m = 0; m < nq; ++m) acc += (double)q[m] * S1[m] * dq * sin((double)q[m] * rr);
    out[k] = (float)((2.0 / 3.14159265358979323846) * acc);
}

// ------------------------------------------------------------------ host helpers
static inline Lattice make_lattice(const float *basis)
{
    Lattice L;
    for (int i = 0; i < 9; ++i) L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    return L;
}

GridParams make_grid(float rmin, float rmax, float bin, int hs)
{
    GridParams g;
    g.rmin = rmin; g.rmax = rmax; g.bin = bin; g.hs = hs;
    g.t2min = sqrt_threshold(rmin);
    g.t2max = sqrt_threshold(rmax);
    g.spill = g_edge_spill;
    g.pad = 0;
    return g;
}

static int check_hist_args(const void *coords_or_dist, int64_t n, const int32_t *mol, const int32_t *el, int nEl,
                           int hs, const float *hintra, const float *hinter)
{
    FRMC_REQUIRE(n >= 0, FRMC_EINVAL, "negative atom count");
    FRMC_REQUIRE(n == 0 || (coords_or_dist && mol && el), FRMC_EINVAL, "NULL input array");
    FRMC_REQUIRE(nEl >= 1 && nEl <= FRMC_MAX_ELEMENTS, FRMC_ELIMIT, "numberOfElements %d outside 1..%d", nEl, FRMC_MAX_ELEMENTS);
    FRMC_REQUIRE(hs >= 1, FRMC_EINVAL, "histSize must be >= 1");
    FRMC_REQUIRE(hintra && hinter, FRMC_EINVAL, "NULL histogram output");
    FRMC_REQUIRE((int64_t)nEl * nEl * hs < (1ll << 31), FRMC_ELIMIT, "histogram too large");
    return FRMC_OK;
}

static int check_elements(const int32_t *el, int64_t n, int nEl)
{
    for (int64_t i = 0; i < n; ++i)
        FRMC_REQUIRE(el[i] >= 0 && el[i] < nEl, FRMC_EINVAL, "elementIndex[%lld]=%d outside 0..%d", (long long)i, el[i], nEl - 1);
    return FRMC_OK;
}

}  // namespace frmc

using namespace frmc;

// Extensions/boundary_conditions_collection.pyx:88-110 transform_coordinates: out[i] = coords[i] . transMatrix, every
// product and sum rounded to float32 in the reference's order (what Engine.py:3223 turns moved real coordinates into
// box coordinates with)
__global__ void transform_coordinates_kernel(const float *__restrict__ coords, int64_t n, const float *__restrict__ m, float *__restrict__ out)
{
    __shared__ float sm[9];
    if (threadIdx.x < 9) sm[threadIdx.x] = m[threadIdx.x];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ox, oy, oz;
    transform_point(sm, coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], ox, oy, oz);
    out[3 * i] = ox; out[3 * i + 1] = oy; out[3 * i + 2] = oz;
}

extern "C" {

int frmc_points_to_coords(int dev, const float *points, const int32_t *from_index, const int64_t *start, int64_t k,
                          const float *coords, int64_t n, const float *basis, int isPBC, int ibc_sign,
                          int want_diff, float *out)
{
    FRMC_REQUIRE(k >= 0 && n >= 0, FRMC_EINVAL, "negative size");
    FRMC_REQUIRE(k <= 65535, FRMC_ELIMIT, "more than 65535 points in one call");
    if (k == 0 || n == 0) return FRMC_OK;
    FRMC_REQUIRE(coords && out && (points || from_index), FRMC_EINVAL, "NULL argument");
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    std::vector<float> pts((size_t)k * 3);
    for (int64_t t = 0; t < k; ++t) {
        if (from_index) {
            FRMC_REQUIRE(from_index[t] >= 0 && from_index[t] < n, FRMC_EINVAL, "index %d outside the coordinates array", from_index[t]);
            for (int d = 0; d < 3; ++d) pts[3 * t + d] = coords[3 * (int64_t)from_index[t] + d];
        } else {
            for (int d = 0; d < 3; ++d) pts[3 * t + d] = points[3 * t + d];
        }
    }
    size_t out_elems = (size_t)n * k * (want_diff ? 3 : 1);
    float *d_coords = (float *)ctx_buffer(c, 0, sizeof(float) * 3 * n);
    float *d_pts = (float *)ctx_buffer(c, 1, sizeof(float) * 3 * k);
    long long *d_start = (long long *)ctx_buffer(c, 2, sizeof(long long) * k);
    float *d_out = (float *)ctx_buffer(c, 3, sizeof(float) * out_elems);
    if (!d_coords || !d_pts || !d_start || !d_out) return FRMC_ENOMEM;
    FRMC_CUDA(cudaMemcpyAsync(d_coords, coords, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_pts, pts.data(), sizeof(float) * 3 * k, cudaMemcpyHostToDevice, c->stream));
    if (start) FRMC_CUDA(cudaMemcpyAsync(d_start, start, sizeof(long long) * k, cudaMemcpyHostToDevice, c->stream));
    Lattice L = make_lattice(basis);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)k);
    const long long *ds = start ? d_start : nullptr;
    if (isPBC) {
        if (want_diff) points_to_coords_kernel<true, true><<<grid, 256, 0, c->stream>>>(d_pts, ds, (int)k, d_coords, n, L, ibc_sign, d_out);
        else points_to_coords_kernel<true, false><<<grid, 256, 0, c->stream>>>(d_pts, ds, (int)k, d_coords, n, L, ibc_sign, d_out);
    } else {
        if (want_diff) points_to_coords_kernel<false, true><<<grid, 256, 0, c->stream>>>(d_pts, ds, (int)k, d_coords, n, L, ibc_sign, d_out);
        else points_to_coords_kernel<false, false><<<grid, 256, 0, c->stream>>>(d_pts, ds, (int)k, d_coords, n, L, ibc_sign, d_out);
    }
    FRMC_LAUNCH_CHECK();
    FRMC_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * out_elems, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    return FRMC_OK;
}

int frmc_from_to_points_differences(int dev, const float *points_from, const float *points_to, int64_t n,
                                    const float *basis, int isPBC, float *out)
{
    FRMC_REQUIRE(n >= 0, FRMC_EINVAL, "negative size");
    if (n == 0) return FRMC_OK;
    FRMC_REQUIRE(points_from && points_to && out, FRMC_EINVAL, "NULL argument");
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    float *d_from = (float *)ctx_buffer(c, 0, sizeof(float) * 3 * n);
    float *d_to = (float *)ctx_buffer(c, 1, sizeof(float) * 3 * n);
    float *d_out = (float *)ctx_buffer(c, 3, sizeof(float) * 3 * n);
    if (!d_from || !d_to || !d_out) return FRMC_ENOMEM;
    FRMC_CUDA(cudaMemcpyAsync(d_from, points_from, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_to, points_to, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    Lattice L = make_lattice(basis);
    unsigned grid = (unsigned)((n + 255) / 256);
    if (isPBC) from_to_kernel<true><<<grid, 256, 0, c->stream>>>(d_from, d_to, n, L, d_out);
    else from_to_kernel<false><<<grid, 256, 0, c->stream>>>(d_from, d_to, n, L, d_out);
    FRMC_LAUNCH_CHECK();
    FRMC_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    return FRMC_OK;
}

int frmc_transform_coordinates(int dev, const float *trans_matrix, const float *coords, int64_t n, float *out)
{
    FRMC_REQUIRE(n >= 0, FRMC_EINVAL, "negative size");
    if (n == 0) return FRMC_OK;
    FRMC_REQUIRE(trans_matrix && coords && out, FRMC_EINVAL, "NULL argument");
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    float *d_in = (float *)ctx_buffer(c, 0, sizeof(float) * 3 * n);
    float *d_m = (float *)ctx_buffer(c, 1, sizeof(float) * 9);
    float *d_out = (float *)ctx_buffer(c, 3, sizeof(float) * 3 * n);
    if (!d_in || !d_m || !d_out) return FRMC_ENOMEM;
    FRMC_CUDA(cudaMemcpyAsync(d_in, coords, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_m, trans_matrix, sizeof(float) * 9, cudaMemcpyHostToDevice, c->stream));
    transform_coordinates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_in, n, d_m, d_out);
    FRMC_LAUNCH_CHECK();
    FRMC_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    return FRMC_OK;
}

int frmc_multiple_pairs_histograms_coords(int dev, const int32_t *indexes, int64_t k, const float *coords, int64_t n,
                                          const float *basis, int isPBC, const int32_t *mol, const int32_t *el,
                                          int nEl, float rmin, float rmax, float bin, int hs, int allAtoms,
                                          float *hintra, float *hinter, uint64_t *edge_overflow)
{
    int rc = check_hist_args(coords, n, mol, el, nEl, hs, hintra, hinter);
    if (rc) return rc;
    FRMC_REQUIRE(k >= 0 && (k == 0 || indexes), FRMC_EINVAL, "bad indexes");
    FRMC_REQUIRE(k < (1ll << 31) && n < (1ll << 31), FRMC_ELIMIT, "more than 2^31 atoms");
    if ((rc = check_elements(el, n, nEl))) return rc;
    for (int64_t t = 0; t < k; ++t)
        FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < n, FRMC_EINVAL, "indexes[%lld]=%d outside 0..%lld", (long long)t, indexes[t], (long long)n - 1);
    const int64_t cells = (int64_t)nEl * nEl * hs;
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    if (k == 0 || n == 0) {
        memset(hintra, 0, sizeof(float) * cells);
        memset(hinter, 0, sizeof(float) * cells);
        if (edge_overflow) *edge_overflow = 0;
        return FRMC_OK;
    }
    float *d_coords = (float *)ctx_buffer(c, 0, sizeof(float) * 3 * n);
    int *d_mol = (int *)ctx_buffer(c, 1, sizeof(int) * n);
    int *d_el = (int *)ctx_buffer(c, 2, sizeof(int) * n);
    int *d_idx = (int *)ctx_buffer(c, 3, sizeof(int) * k);
    unsigned int *d_counts = (unsigned int *)ctx_buffer(c, 4, sizeof(unsigned int) * 2 * cells + 16);
    float *d_out = (float *)ctx_buffer(c, 5, sizeof(float) * 2 * cells);
    if (!d_coords || !d_mol || !d_el || !d_idx || !d_counts || !d_out) return FRMC_ENOMEM;
    unsigned long long *d_ov = (unsigned long long *)(d_counts + 2 * cells + (2 * cells) % 2);
    FRMC_CUDA(cudaMemcpyAsync(d_coords, coords, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_mol, mol, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_el, el, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_idx, indexes, sizeof(int) * k, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned int) * 2 * cells + 16, c->stream));
    Lattice L = make_lattice(basis);
    GridParams g = make_grid(rmin, rmax, bin, hs);
    int mode = choose_mode(L.b, isPBC, coords, n);
    int chunks = (int)((k + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK);
    FRMC_REQUIRE(chunks <= 65535, FRMC_ELIMIT, "too many rows for one call (%lld)", (long long)k);
    long long want = (n + 255) / 256;
    long long cap = (long long)c->sm_count * 8;
    dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)chunks);
#define LAUNCH_ROWS(M) rows_hist_kernel<M><<<grid, 256, 0, c->stream>>>(d_coords, d_mol, d_el, n, d_idx, (int)k, allAtoms, L, g, nEl, d_counts, d_ov)
    switch (mode) {
        case MODE_IBC: LAUNCH_ROWS(MODE_IBC); break;
        case MODE_ORTHO_FAST: LAUNCH_ROWS(MODE_ORTHO_FAST); break;
        case MODE_TRI_FAST: LAUNCH_ROWS(MODE_TRI_FAST); break;
        case MODE_ORTHO_GEN: LAUNCH_ROWS(MODE_ORTHO_GEN); break;
        default: LAUNCH_ROWS(MODE_TRI_GEN); break;
    }
#undef LAUNCH_ROWS
    FRMC_LAUNCH_CHECK();
    counts_to_float_kernel<<<(unsigned)((2 * cells + 255) / 256), 256, 0, c->stream>>>(d_counts, nullptr, d_out, 2 * cells);
    FRMC_LAUNCH_CHECK();
    unsigned long long ov = 0;
    FRMC_CUDA(cudaMemcpyAsync(hintra, d_out, sizeof(float) * cells, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(hinter, d_out + cells, sizeof(float) * cells, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(&ov, d_ov, sizeof(ov), cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    if (edge_overflow) *edge_overflow = ov;
    return FRMC_OK;
}

static int dists_hist_common(int dev, const int32_t *indexes, int64_t k, const float *distances, int64_t n,
                             int64_t stride_row, int64_t stride_col, int64_t dist_elems, const int32_t *mol,
                             const int32_t *el, int nEl, float rmin, float rmax, float bin, int hs, int allAtoms,
                             float *hintra, float *hinter, int in_place, uint64_t *edge_overflow)
{
    int rc = check_hist_args(distances, n, mol, el, nEl, hs, hintra, hinter);
    if (rc) return rc;
    FRMC_REQUIRE(k >= 0 && k <= 65535 && (k == 0 || indexes), FRMC_EINVAL, "bad indexes (k=%lld)", (long long)k);
    if ((rc = check_elements(el, n, nEl))) return rc;
    for (int64_t t = 0; t < k; ++t)
        FRMC_REQUIRE(indexes[t] >= 0 && indexes[t] < n, FRMC_EINVAL, "indexes[%lld]=%d outside 0..%lld", (long long)t, indexes[t], (long long)n - 1);
    const int64_t cells = (int64_t)nEl * nEl * hs;
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    if (k == 0 || n == 0) {
        if (!in_place) { memset(hintra, 0, sizeof(float) * cells); memset(hinter, 0, sizeof(float) * cells); }
        if (edge_overflow) *edge_overflow = 0;
        return FRMC_OK;
    }
    float *d_dist = (float *)ctx_buffer(c, 0, sizeof(float) * dist_elems);
    int *d_mol = (int *)ctx_buffer(c, 1, sizeof(int) * n);
    int *d_el = (int *)ctx_buffer(c, 2, sizeof(int) * n);
    int *d_idx = (int *)ctx_buffer(c, 3, sizeof(int) * k);
    unsigned int *d_counts = (unsigned int *)ctx_buffer(c, 4, sizeof(unsigned int) * 2 * cells + 16);
    float *d_out = (float *)ctx_buffer(c, 5, sizeof(float) * 2 * cells);
    float *d_base = (float *)ctx_buffer(c, 6, sizeof(float) * 2 * cells);
    if (!d_dist || !d_mol || !d_el || !d_idx || !d_counts || !d_out || !d_base) return FRMC_ENOMEM;
    unsigned long long *d_ov = (unsigned long long *)(d_counts + 2 * cells + (2 * cells) % 2);
    FRMC_CUDA(cudaMemcpyAsync(d_dist, distances, sizeof(float) * dist_elems, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_mol, mol, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_el, el, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_idx, indexes, sizeof(int) * k, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned int) * 2 * cells + 16, c->stream));
    if (in_place) {
        FRMC_CUDA(cudaMemcpyAsync(d_base, hintra, sizeof(float) * cells, cudaMemcpyHostToDevice, c->stream));
        FRMC_CUDA(cudaMemcpyAsync(d_base + cells, hinter, sizeof(float) * cells, cudaMemcpyHostToDevice, c->stream));
    }
    GridParams g = make_grid(rmin, rmax, bin, hs);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)k);
    dists_hist_kernel<<<grid, 256, 0, c->stream>>>(d_dist, n, stride_row, stride_col, d_mol, d_el, d_idx, (int)k, allAtoms, g, nEl, d_counts, d_ov);
    FRMC_LAUNCH_CHECK();
    counts_to_float_kernel<<<(unsigned)((2 * cells + 255) / 256), 256, 0, c->stream>>>(d_counts, in_place ? d_base : nullptr, d_out, 2 * cells);
    FRMC_LAUNCH_CHECK();
    unsigned long long ov = 0;
    FRMC_CUDA(cudaMemcpyAsync(hintra, d_out, sizeof(float) * cells, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(hinter, d_out + cells, sizeof(float) * cells, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(&ov, d_ov, sizeof(ov), cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    if (edge_overflow) *edge_overflow = ov;
    return FRMC_OK;
}

int frmc_multiple_pairs_histograms_dists(int dev, const int32_t *indexes, int64_t k, const float *distances, int64_t n,
                                         const int32_t *mol, const int32_t *el, int nEl, float rmin, float rmax,
                                         float bin, int hs, int allAtoms, float *hintra, float *hinter,
                                         uint64_t *edge_overflow)
{
    return dists_hist_common(dev, indexes, k, distances, n, k, 1, n * k, mol, el, nEl, rmin, rmax, bin, hs, allAtoms,
                             hintra, hinter, 0, edge_overflow);
}

int frmc_single_pairs_histograms(int dev, int32_t atomIndex, const float *distances, int64_t dstride, int64_t n,
                                 const int32_t *mol, const int32_t *el, int nEl, int hs, float *hintra,
                                 float *hinter, float rmin, float rmax, float bin, int allAtoms,
                                 uint64_t *edge_overflow)
{
    FRMC_REQUIRE(dstride >= 1, FRMC_EINVAL, "distance stride must be >= 1");
    int32_t idx = atomIndex;
    int64_t elems = n > 0 ? (n - 1) * dstride + 1 : 0;
    return dists_hist_common(dev, &idx, 1, distances, n, dstride, 0, elems, mol, el, nEl, rmin, rmax, bin, hs, allAtoms,
                             hintra, hinter, 1, edge_overflow);
}

static int reciprocal_common(int dev, int which, const float *a, const float *b, int64_t n, const float *qs, int64_t m,
                             float rho, float *out, int64_t n_out)
{
    FRMC_REQUIRE(n >= 2 && m >= 1, FRMC_EINVAL, "need at least 2 abscissa points and 1 output point");
    FRMC_REQUIRE(a && b && qs && out, FRMC_EINVAL, "NULL argument");
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    float *d_a = (float *)ctx_buffer(c, 0, sizeof(float) * n);
    float *d_b = (float *)ctx_buffer(c, 1, sizeof(float) * n);
    float *d_q = (float *)ctx_buffer(c, 2, sizeof(float) * m);
    float *d_out = (float *)ctx_buffer(c, 3, sizeof(float) * n_out);
    if (!d_a || !d_b || !d_q || !d_out) return FRMC_ENOMEM;
    FRMC_CUDA(cudaMemcpyAsync(d_a, a, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_b, b, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    FRMC_CUDA(cudaMemcpyAsync(d_q, qs, sizeof(float) * m, cudaMemcpyHostToDevice, c->stream));
    unsigned grid = (unsigned)((n_out + 63) / 64);
    if (which == 0) r_to_q_kernel<false><<<grid, 64, 0, c->stream>>>(d_a, d_b, n, d_q, m, 0.f, d_out);
    else if (which == 1) {
        float fact = 4.0f * 3.1415927f * rho;   // FLOAT32_FOUR * FLOAT32_PI * rho, fp32 (reciprocal_space.pyx:64)
        r_to_q_kernel<true><<<grid, 64, 0, c->stream>>>(d_a, d_b, n, d_q, m, fact, d_out);
    } else {
        // a = q values [n], qs = r values [m], b = sq [n]
        q_to_r_kernel<<<grid, 64, 0, c->stream>>>(d_a, d_q, d_b, n, m, d_out);
    }
    FRMC_LAUNCH_CHECK();
    FRMC_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * n_out, cudaMemcpyDeviceToHost, c->stream));
    FRMC_CUDA(cudaStreamSynchronize(c->stream));
    return FRMC_OK;
}

int frmc_shape_function(int dev, const float *hintra, const float *hinter, int nEl, int hs, int n_pairs, const int32_t *pair_a,
                        const int32_t *pair_b, const double *pair_coef, const float *shell_volumes, const float *shell_centers,
                        double rho0, const float *q, int nq, const float *r, int nr, float *out)
{
    FRMC_REQUIRE(hintra && hinter && pair_a && pair_b && pair_coef && shell_volumes && shell_centers && q && r && out, FRMC_EINVAL, "NULL argument");
    FRMC_REQUIRE(nEl >= 1 && nEl <= FRMC_MAX_ELEMENTS && hs >= 2 && nq >= 2 && nr >= 1 && n_pairs >= 1, FRMC_EINVAL, "bad sizes");
    DeviceCtx *c = get_ctx(dev);
    if (!c) return FRMC_ECUDA;
    const size_t cells = (size_t)nEl * nEl * hs;
    float *d_h = (float *)ctx_buffer(c, 0, sizeof(float) * 2 * cells);
    float *d_f = (float *)ctx_buffer(c, 1, sizeof(float) * ((size_t)2 * hs + nq + 2 * (size_t)nr));
    double *d_d = (double *)ctx_buffer(c, 2, sizeof(double) * ((size_t)n_pairs + hs + nq) + sizeof(int) * 2 * (size_t)n_pairs);
    if (!d_h || !d_f || !d_d) return FRMC_ENOMEM;
    float *d_vol = d_f, *d_cen = d_f + hs, *d_q = d_f + 2 * hs, *d_r = d_q + nq, *d_out = d_r + nr;
    double *d_coef = d_d, *d_G = d_d + n_pairs, *d_S1 = d_G + hs;
    int *d_pa = reinterpret_cast<int *>(d_S1 + nq), *d_pb = d_pa + n_pairs;
    cudaStream_t st = c->stream;
    FRMC_CUDA(cudaMemcpyAsync(d_h, hintra, sizeof(float) * cells, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_h + cells, hinter, sizeof(float) * cells, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_vol, shell_volumes, sizeof(float) * hs, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_cen, shell_centers, sizeof(float) * hs, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_q, q, sizeof(float) * nq, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_r, r, sizeof(float) * nr, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_coef, pair_coef, sizeof(double) * n_pairs, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_pa, pair_a, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, st));
    FRMC_CUDA(cudaMemcpyAsync(d_pb, pair_b, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, st));
    shape_gr_kernel<<<(hs + 127) / 128, 128, 0, st>>>(d_h, d_h + cells, nEl, hs, n_pairs, d_pa, d_pb, d_coef, d_vol, d_cen, rho0, d_G);
    FRMC_LAUNCH_CHECK();
    shape_sq_kernel<<<(nq + 63) / 64, 64, 0, st>>>(d_G, d_cen, hs, d_q, nq, d_S1);
    FRMC_LAUNCH_CHECK();
    shape_back_kernel<<<(nr + 63) / 64, 64, 0, st>>>(d_S1, d_q, nq, d_r, nr, d_out);
    FRMC_LAUNCH_CHECK();
    FRMC_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * nr, cudaMemcpyDeviceToHost, st));
    FRMC_CUDA(cudaStreamSynchronize(st));
    return FRMC_OK;
}

int frmc_Gr_to_sq(int dev, const float *distances, const float *Gr, int64_t n, const float *qrange, int64_t m, float *sq)
{
    return reciprocal_common(dev, 0, distances, Gr, n, qrange, m, 0.f, sq, m);
}

int frmc_gr_to_sq(int dev, const float *distances, const float *gr, int64_t n, const float *qrange, int64_t m, float rho, float *sq)
{
    return reciprocal_common(dev, 1, distances, gr, n, qrange, m, rho, sq, m);
}

int frmc_sq_to_Gr(int dev, const float *qvalues, const float *rvalues, const float *sq, int64_t m, int64_t n, float *Gr)
{
    // m q-points, n r-points
    return reciprocal_common(dev, 2, qvalues, sq, m, rvalues, n, 0.f, Gr, n);
}

}  // extern "C"
