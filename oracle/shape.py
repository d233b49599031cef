"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's ``ShapeFunction``
(Constraints/Collection.py:20-125) as used by ``PairDistributionConstraint._update_shape_array``
(Constraints/PairDistributionConstraints.py:316-343).  Nothing under ``fullrmc_b200/`` may import this; the product
computes the same quantity on the device (fullrmc_b200/shape.py -> frmc_shape_function).

Parity status: PINNED -- tests/test_golden_constraints.py replays the SiOx trajectory of the unmodified reference
class (six refreshes) and requires these arrays to equal the reference's bit for bit.  It follows the reference's
numpy expressions one for one:

* the r-grid and Q values of the private StructureFactorConstraint (StructureFactorConstraints.py:330-350),
* its *plotting-path* total ``get_constraint_value()["total"]`` (StructureFactorConstraints.py:836-896; note the
  operation order differs from the fitting path ``__get_total_Sq``),
* the sine back-transform ``G(r) = 2/pi * sum_q q (S(q)-1) sin(q r) dq`` (Collection.py:83-91).
"""
import numpy as np

from .epilogue import elements_pairs, gr2sq_matrix, shell_arrays_from_edges

FLOAT_TYPE = np.float32
PI = FLOAT_TYPE(np.pi)


def auto_rmax(isPBC, basisVectors, realCoordinates):
    """rmax when the parameters leave it open (PairDistributionConstraints.py:323-334)"""
    if isPBC:
        lengths = [np.linalg.norm(v) for v in np.asarray(basisVectors)]        # boundaryConditions.get_a/b/c
        return FLOAT_TYPE(np.max(lengths) + 10)
    real = np.asarray(realCoordinates)
    coordsCenter = np.sum(real, axis=0) / real.shape[0]
    coordinates = real - coordsCenter
    distances = np.sqrt(np.sum(coordinates ** 2, axis=1))
    maxDistance = 2. * np.max(distances)
    return FLOAT_TYPE(maxDistance + 10)


def shape_grid(rmin, rmax, dr):
    """edges, centres, volumes of the private StructureFactorConstraint (rmax given: :340-350)"""
    rmin, rmax, dr = FLOAT_TYPE(rmin), FLOAT_TYPE(rmax), FLOAT_TYPE(dr)
    edges = np.arange(rmin, rmax + dr, dr).astype(FLOAT_TYPE)
    centers = (edges[0:-1] + edges[1:]) / FLOAT_TYPE(2.)
    return edges, centers, shell_arrays_from_edges(edges)


def plotting_total_Sq(intra, inter, elements, n_per_element, weighting, volume, numberOfAtoms, shell_centers,
                      shell_volumes, gr2sq):
    """StructureFactorConstraint._get_constraint_value(...)["total"] with scale factor 1, no window (:836-896)"""
    volume = FLOAT_TYPE(volume)
    gr = np.zeros(shell_centers.shape[0], dtype=FLOAT_TYPE)
    for pair in elements_pairs(elements):
        wij = weighting.get(pair[0] + "-" + pair[1], None)
        if wij is None:
            wij = weighting[pair[1] + "-" + pair[0]]
        ni, nj = n_per_element[pair[0]], n_per_element[pair[1]]
        idi, idj = elements.index(pair[0]), elements.index(pair[1])
        sf_intra = np.zeros(shell_centers.shape[0], dtype=FLOAT_TYPE)
        sf_inter = np.zeros(shell_centers.shape[0], dtype=FLOAT_TYPE)
        if idi == idj:
            Nij = ni * (ni - 1) / 2.0
            sf_intra += intra[idi, idj, :]
            sf_inter += inter[idi, idj, :]
        else:
            Nij = ni * nj
            sf_intra += intra[idi, idj, :] + intra[idj, idi, :]
            sf_inter += inter[idi, idj, :] + inter[idj, idi, :]
        nij = sf_intra + sf_inter
        dij = nij / shell_volumes
        Dij = Nij / volume
        gr += wij * dij / Dij
    rho0 = FLOAT_TYPE(numberOfAtoms / volume)
    Gr = (FLOAT_TYPE(4.) * PI * shell_centers * rho0) * (gr - 1)
    return np.sum(Gr.reshape((-1, 1)) * gr2sq, axis=0) + 1


def Gr_from_Sq(qValues, rValues, Sq):
    """ShapeFunction.__get_Gr_from_Sq (Collection.py:83-91)"""
    Gr = np.zeros(len(rValues), dtype=FLOAT_TYPE)
    sq_1 = Sq - 1
    qsq_1 = qValues * sq_1
    dq = qValues[1] - qValues[0]
    for ridx, r in enumerate(rValues):
        sinqr_dq = dq * np.sin(qValues * r)
        Gr[ridx] = (2. / PI) * np.sum(qsq_1 * sinqr_dq)
    return Gr


def get_Gr_shape_function(rValues, boxCoordinates, basisVectors, isPBC, moleculesIndex, elementsIndex, elements,
                          numberOfAtomsPerElement, volume, weighting, qmin=0.001, qmax=1, dq=0.005, rmin=0.00, rmax=100, dr=1,
                          full_histogram=None):
    """ShapeFunction(engine, ...).get_Gr_shape_function(rValues) for the given engine arrays.

    ``weighting`` is the private constraint's weighting scheme ("A-B" -> float32), i.e. what
    ``get_normalized_weighting`` returns for the "atomicNumber" property; ``full_histogram`` defaults to the
    C oracle's ``full_pairs_histograms_coords``."""
    if full_histogram is None:
        from .pairhist import full_pairs_histograms_coords as full_histogram
    elements = list(elements)
    Q = np.arange(FLOAT_TYPE(qmin), FLOAT_TYPE(qmax), FLOAT_TYPE(dq))
    qValues = np.transpose([Q, np.zeros(len(Q))]).astype(FLOAT_TYPE)[:, 0]
    edges, centers, volumes = shape_grid(rmin, rmax, dr)
    hs = len(edges) - 1
    intra, inter = full_histogram(boxCoords=np.ascontiguousarray(boxCoordinates, dtype=FLOAT_TYPE),
                                  basis=np.ascontiguousarray(basisVectors, dtype=FLOAT_TYPE), isPBC=bool(isPBC),
                                  moleculeIndex=np.ascontiguousarray(moleculesIndex, dtype=np.int32),
                                  elementIndex=np.ascontiguousarray(elementsIndex, dtype=np.int32),
                                  numberOfElements=len(elements), minDistance=edges[0], maxDistance=edges[-1],
                                  bin=FLOAT_TYPE(dr), histSize=hs)
    Sq = plotting_total_Sq(intra, inter, elements, numberOfAtomsPerElement, weighting, volume, len(elementsIndex), centers,
                           volumes, gr2sq_matrix(qValues, centers))
    return Gr_from_Sq(qValues, np.asarray(rValues, dtype=FLOAT_TYPE), Sq)
