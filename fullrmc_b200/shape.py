"""Shape function of a finite (nano-particle) system on the device: what the reference's ``ShapeFunction``
(Constraints/Collection.py:20-125) hands ``PairDistributionConstraint._update_shape_array``
(Constraints/PairDistributionConstraints.py:316-343).

Both steps run on the GPU: the full pair histogram of the current configuration on the shape function's own coarse
r-grid (``Core.pairs_histograms.full_pairs_histograms_coords``) and the chain histogram -> G(r) -> S(q) - 1 -> sine
back-transform onto the constraint's r values (``frmc_shape_function``, csrc/stateless.cu).  This module only lays
out the inputs: the r-grid and q values the reference's private StructureFactorConstraint would build from
(rmin, rmax, dr) and (qmin, qmax, dq), and one coefficient per element pair.
"""
import ctypes
import itertools

import numpy as np

from . import _lib as L
from .model import FLOAT_TYPE, shell_volumes_from_edges


def default_rmax(isPBC, basisVectors, coordinates):
    """the histogram's reach when the parameters leave rmax open: ten Angstrom beyond the longest cell edge of a
    periodic box, beyond the diameter about the centroid of a finite one (PairDistributionConstraints.py:323-334)"""
    if isPBC:
        return FLOAT_TYPE(max(float(np.linalg.norm(v)) for v in np.asarray(basisVectors)) + 10)
    xyz = np.asarray(coordinates)
    radius = np.sqrt(((xyz - xyz.sum(axis=0) / xyz.shape[0]) ** 2).sum(axis=1)).max()
    return FLOAT_TYPE(2. * radius + 10)


def get_Gr_shape_function(rValues, boxCoordinates, basisVectors, isPBC, moleculesIndex, elementsIndex, elements,
                          numberOfAtomsPerElement, volume, weighting, qmin=0.001, qmax=1, dq=0.005, rmin=0.00, rmax=100, dr=1):
    """G_shape(r) on ``rValues`` for the given engine arrays (ShapeFunction(engine, ...).get_Gr_shape_function).
    ``weighting`` is the weighting scheme of the private constraint ("A-B" -> float32: what
    ``get_normalized_weighting`` returns for the "atomicNumber" property)."""
    from .Core.pairs_histograms import full_pairs_histograms_coords
    lib = L.load_library()
    elements = list(elements)
    nEl = len(elements)
    q = np.ascontiguousarray(np.arange(FLOAT_TYPE(qmin), FLOAT_TYPE(qmax), FLOAT_TYPE(dq)), dtype=FLOAT_TYPE)
    step = FLOAT_TYPE(dr)
    edges = np.arange(FLOAT_TYPE(rmin), FLOAT_TYPE(rmax) + step, step).astype(FLOAT_TYPE)      # StructureFactorConstraints.py:341
    centers = np.ascontiguousarray((edges[:-1] + edges[1:]) / FLOAT_TYPE(2.), dtype=FLOAT_TYPE)
    volumes = np.ascontiguousarray(shell_volumes_from_edges(edges), dtype=FLOAT_TYPE)
    hs = centers.shape[0]
    intra, inter = full_pairs_histograms_coords(boxCoords=np.ascontiguousarray(boxCoordinates, dtype=FLOAT_TYPE),
                                                basis=np.ascontiguousarray(basisVectors, dtype=FLOAT_TYPE), isPBC=bool(isPBC),
                                                moleculeIndex=np.ascontiguousarray(moleculesIndex, dtype=np.int32),
                                                elementIndex=np.ascontiguousarray(elementsIndex, dtype=np.int32),
                                                numberOfElements=nEl, minDistance=edges[0], maxDistance=edges[-1], bin=step, histSize=hs)
    pairs = sorted(itertools.combinations_with_replacement(elements, 2))
    pa = np.array([elements.index(a) for a, _ in pairs], np.int32)
    pb = np.array([elements.index(b) for _, b in pairs], np.int32)
    coef = np.empty(len(pairs), np.float64)                       # w_ij / D_ij, D_ij = (pairs of the kind) / volume
    for k, (a, b) in enumerate(pairs):
        w = weighting.get(a + "-" + b, weighting.get(b + "-" + a))
        na, nb = numberOfAtomsPerElement[a], numberOfAtomsPerElement[b]
        kinds = na * (na - 1) / 2.0 if a == b else float(na) * nb
        coef[k] = float(w) * float(volume) / kinds if kinds > 0 else 0.0
    r = np.ascontiguousarray(rValues, dtype=FLOAT_TYPE)
    out = np.empty(r.shape[0], FLOAT_TYPE)
    rho0 = float(len(elementsIndex)) / float(volume)
    rc = lib.frmc_shape_function(L.device_index(), L.ptr(intra, L.c_f32p), L.ptr(inter, L.c_f32p), nEl, hs, len(pairs), L.ptr(pa, L.c_i32p),
                                 L.ptr(pb, L.c_i32p), coef.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), L.ptr(volumes, L.c_f32p),
                                 L.ptr(centers, L.c_f32p), rho0, L.ptr(q, L.c_f32p), q.shape[0], L.ptr(r, L.c_f32p), r.shape[0],
                                 L.ptr(out, L.c_f32p))
    L.check(rc, "shape_function")
    return out
