"""numpy restatement of the constraint-level math of the three hot-path constraints.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, expression by expression and in float32 like the reference (numpy >= 2 scalar
promotion, i.e. NEP 50: Python scalars are weak, np.float32 scalars stay float32):

* PairDistributionConstraint.__get_total_Gr      Constraints/PairDistributionConstraints.py:847-895
* PairCorrelationConstraint.__get_total_gr       Constraints/PairCorrelationConstraints.py:126-169
* StructureFactorConstraint.__get_total_Sq       Constraints/StructureFactorConstraints.py:780-822
  (+ _get_Sq_from_Gr :772-773, _apply_scale_factor :775-778, Reduced variant :1253-1260,
   __set_Gr_2_Sq_matrix :302-312)
* compute_standard_error                         PairDistributionConstraints.py:810-838
* fit_scale_factor                               Core/Constraint.py:1363-1395
* the per-move sequence compute_before_move / compute_after_move / accept_move
                                                 PairDistributionConstraints.py:1044-1152

Window-function convolution and the multiframe prior (unused by the five BASELINE
configs) are not restated.

Parity status: pinned against the reference's real constraint classes run under
third-party stubs (tests/gen_golden.py -> tests/golden/*.npz).
"""
import itertools

import numpy as np

FLOAT_TYPE = np.float32
PI = FLOAT_TYPE(np.pi)            # Globals.py:45


def elements_pairs(elements):
    """sorted(itertools.combinations_with_replacement(elements, 2)) -- PairDistributionConstraints.py:485"""
    return sorted(itertools.combinations_with_replacement(list(elements), 2))


def normalized_weighting(numbers, weights):
    """Restatement of pdbparser.Utilities.Collection.get_normalized_weighting (pdbparser >= 0.1.8,
    not vendored under /root/reference: PARITY UNPINNED for these scalar inputs).  Faber-Ziman:
    w_ij = c_i c_j b_i b_j / (sum_k c_k b_k)^2, doubled for i != j; keys "A-B"."""
    els = list(numbers.keys())
    total = float(sum(numbers.values()))
    c = {e: numbers[e] / total for e in els}
    norm = sum(c[e] * float(weights[e]) for e in els) ** 2
    out = {}
    for i, a in enumerate(els):
        for b in els[i:]:
            w = c[a] * c[b] * float(weights[a]) * float(weights[b]) / norm
            if a != b:
                w *= 2.0
            out[a + "-" + b] = w
    return out


def shell_arrays_from_edges(edges):
    """shellCenters / shellVolumes from float32 edges -- PairDistributionConstraints.py:748-758,
    StructureFactorConstraints.py:344-350."""
    edges = np.asarray(edges, dtype=FLOAT_TYPE)
    shell_volumes = FLOAT_TYPE(4.0 / 3.) * PI * ((edges[1:]) ** 3 - edges[0:-1] ** 3)
    return shell_volumes


def gr2sq_matrix(q_values, shell_centers):
    """StructureFactorConstraints.py:302-312 (__set_Gr_2_Sq_matrix)."""
    Qs = np.asarray(q_values, dtype=FLOAT_TYPE)
    Rs = np.asarray(shell_centers, dtype=FLOAT_TYPE)
    dr = Rs[1] - Rs[0]
    qr = Rs.reshape((-1, 1)) * (np.ones((len(Rs), 1), dtype=FLOAT_TYPE) * Qs)
    sinqr = np.sin(qr)
    sinqr_q = sinqr / Qs
    return dr * sinqr_q


def _weighted_sum(intra, inter, elements, n_per_element, weighting, volume, cast_D):
    hs = intra.shape[2]
    acc = np.zeros(hs, dtype=FLOAT_TYPE)
    for pair in elements_pairs(elements):
        wij = weighting.get(pair[0] + "-" + pair[1], None)
        if wij is None:
            wij = weighting[pair[1] + "-" + pair[0]]
        wij = FLOAT_TYPE(wij)
        ni = n_per_element[pair[0]]
        nj = n_per_element[pair[1]]
        idi = elements.index(pair[0])
        idj = elements.index(pair[1])
        if idi == idj:
            Nij = ni * (ni - 1) / 2.0
            Dij = Nij / volume
            if cast_D:
                Dij = FLOAT_TYPE(Dij)
            nij = intra[idi, idj, :] + inter[idi, idj, :]
            acc += wij * nij / Dij
        else:
            Nij = ni * nj
            Dij = Nij / volume
            if cast_D:
                Dij = FLOAT_TYPE(Dij)
            nij = intra[idi, idj, :] + intra[idj, idi, :] + inter[idi, idj, :] + inter[idj, idi, :]
            acc += wij * nij / Dij
    return acc


def total_Gr(intra, inter, elements, n_per_element, weighting, volume, rho0, shell_centers, shell_volumes,
             shape_array=None, scale_factor=1.0, refit=None):
    """PairDistributionConstraint.__get_total_Gr (PairDistributionConstraints.py:847-895).
    refit = (experimental, data_weights, sf_min, sf_max): this evaluation refits the scale factor
    (get_adjusted_scale_factor, Core/Constraint.py:1397-1423) and the function returns (total, SF)."""
    volume, rho0 = FLOAT_TYPE(volume), FLOAT_TYPE(rho0)
    Gr = _weighted_sum(intra, inter, elements, n_per_element, weighting, volume, cast_D=True)
    Gr /= shell_volumes
    Gr = (4. * PI * shell_centers * rho0) * (Gr - 1)
    if shape_array is not None:
        Gr -= shape_array
    if refit is not None:
        scale_factor = fit_scale_factor(refit[0], Gr, refit[1], refit[2], refit[3])
    if scale_factor != 1:
        Gr *= FLOAT_TYPE(scale_factor)
    return Gr if refit is None else (Gr, FLOAT_TYPE(scale_factor))


def total_gr(intra, inter, elements, n_per_element, weighting, volume, rho0, shell_centers, shell_volumes,
             shape_array=None, scale_factor=1.0, refit=None):
    """PairCorrelationConstraint.__get_total_gr (PairCorrelationConstraints.py:126-169); refit as in total_Gr,
    fitted on G(r) = 4 pi r rho0 (g - 1) (:171-184)."""
    volume, rho0 = FLOAT_TYPE(volume), FLOAT_TYPE(rho0)
    gr = _weighted_sum(intra, inter, elements, n_per_element, weighting, volume, cast_D=False)
    gr /= shell_volumes
    if shape_array is not None:
        gr -= shape_array
    if refit is not None:
        expGr = (4. * PI * shell_centers * rho0) * (refit[0] - 1)
        Gr_ = (4. * PI * shell_centers * rho0) * (gr - 1)
        scale_factor = fit_scale_factor(expGr, Gr_, refit[1], refit[2], refit[3])
    if scale_factor != 1:
        scale_factor = FLOAT_TYPE(scale_factor)
        Gr = (4. * PI * shell_centers * rho0) * (gr - 1)
        Gr *= scale_factor
        gr = 1. + Gr / (4. * PI * shell_centers * rho0)
    return gr if refit is None else (gr, FLOAT_TYPE(scale_factor))


def Sq_from_Gr(Gr, gr2sq, reduced=False):
    """_get_Sq_from_Gr (StructureFactorConstraints.py:772-773; Reduced :1253-1254)."""
    s = np.sum(Gr.reshape((-1, 1)) * gr2sq, axis=0)
    return s if reduced else s + 1


def total_Sq(intra, inter, elements, n_per_element, weighting, volume, rho0, shell_centers, shell_volumes,
             gr2sq, scale_factor=1.0, reduced=False, return_Gr=False, refit=None):
    """StructureFactorConstraint.__get_total_Sq (StructureFactorConstraints.py:780-822); refit as in total_Gr,
    fitted on S(Q)-1 against experimental-1 (:824-834; the reduced constraint inherits that override)."""
    volume, rho0 = FLOAT_TYPE(volume), FLOAT_TYPE(rho0)
    Gr = _weighted_sum(intra, inter, elements, n_per_element, weighting, volume, cast_D=False)
    Gr /= shell_volumes
    Gr = (FLOAT_TYPE(4.) * PI * shell_centers * rho0) * (Gr - 1)
    Sq = Sq_from_Gr(Gr, gr2sq, reduced=reduced)
    if refit is not None:
        scale_factor = fit_scale_factor(refit[0] - 1, Sq - 1, refit[1], refit[2], refit[3])
    if scale_factor != 1:
        scale_factor = FLOAT_TYPE(scale_factor)
        Sq = scale_factor * Sq if reduced else scale_factor * (Sq - 1) + 1
    if refit is not None:
        return Sq, FLOAT_TYPE(scale_factor)
    if return_Gr:
        return Sq, Gr
    return Sq


def apply_prior_and_window(total, prior=None, weight=None, window=None):
    """tail of __get_total_Gr / __get_total_gr / __get_total_Sq after the scale factor: multiframe prior
    (Core/Constraint.py:1160-1177) then window convolution (PairDistributionConstraints.py:890-893)."""
    if weight is not None:
        total = prior + FLOAT_TYPE(weight) * total
    if window is not None:
        total = np.convolve(total, window, 'same')
    return total


def standard_error(experimental, model, data_weights=None):
    """compute_standard_error (PairDistributionConstraints.py:810-838; StructureFactorConstraints.py:742-770)."""
    diff = experimental - model
    if data_weights is None:
        return np.add.reduce((diff) ** 2)
    return np.add.reduce(data_weights * ((diff) ** 2))


def fit_scale_factor(experimental, model, data_weights, sf_min, sf_max):
    """ExperimentalConstraint.fit_scale_factor (Core/Constraint.py:1363-1395)."""
    if data_weights is None:
        SF = FLOAT_TYPE(np.sum(model * experimental) / np.sum(model ** 2))
    else:
        SF = FLOAT_TYPE(np.sum(data_weights * model * experimental) / np.sum(model ** 2))
    SF = max(SF, sf_min)
    SF = min(SF, sf_max)
    return SF


# ---------------------------------------------------------------- explicit summation orders
def numpy_pairwise_sum(a):
    """Scalar restatement of numpy's FLOAT_pairwise_sum (numpy/_core/src/umath/loops_utils.h.src),
    the order np.add.reduce uses on a contiguous float32 vector.  The CUDA chi^2 kernel follows
    this order; tests/test_oracle.py checks it equals np.add.reduce bit for bit."""
    a = np.asarray(a, dtype=FLOAT_TYPE)
    n = a.shape[0]
    if n < 8:
        res = FLOAT_TYPE(0.)
        for i in range(n):
            res = FLOAT_TYPE(res + a[i])
        return res
    if n <= 128:
        r = [FLOAT_TYPE(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = FLOAT_TYPE(r[j] + a[i + j])
            i += 8
        res = FLOAT_TYPE(FLOAT_TYPE(FLOAT_TYPE(r[0] + r[1]) + FLOAT_TYPE(r[2] + r[3])) +
                         FLOAT_TYPE(FLOAT_TYPE(r[4] + r[5]) + FLOAT_TYPE(r[6] + r[7])))
        while i < n:
            res = FLOAT_TYPE(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return FLOAT_TYPE(numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:]))


def sequential_Sq(Gr, gr2sq):
    """Row-sequential fp32 accumulation: acc[m] = (...((G0*M0m) + G1*M1m) + ...), the order numpy
    uses for np.sum(Gr[:,None]*M, axis=0) on a C-contiguous matrix and the order the CUDA S(Q)
    kernel follows."""
    acc = np.zeros(gr2sq.shape[1], dtype=FLOAT_TYPE)
    for r in range(gr2sq.shape[0]):
        acc = acc + Gr[r] * gr2sq[r]
    return acc


# ---------------------------------------------------------------- per-move reference sequence
def move_delta(hist_fns, relative_indexes, box_coords, basis, is_pbc, mol, el, n_el, rmin, rmax, bin, hs):
    """activeAtomsData = M - F for the CURRENT coordinates (compute_before_move,
    PairDistributionConstraints.py:1053-1078).  hist_fns = (multiple_pairs_histograms_coords,
    full_pairs_histograms_coords) from either oracle/_ref or oracle.pairhist."""
    multiple, full = hist_fns
    idx = np.asarray(relative_indexes, dtype=np.int32)
    kw = dict(basis=basis, isPBC=is_pbc, numberOfElements=n_el, minDistance=rmin, maxDistance=rmax, bin=bin,
              histSize=hs)
    intraM, interM = multiple(indexes=idx, boxCoords=box_coords, moleculeIndex=mol, elementIndex=el,
                              allAtoms=True, **kw)
    intraF, interF = full(boxCoords=np.ascontiguousarray(box_coords[idx]), moleculeIndex=np.ascontiguousarray(mol[idx]),
                          elementIndex=np.ascontiguousarray(el[idx]), **kw)
    return intraM - intraF, interM - interF
