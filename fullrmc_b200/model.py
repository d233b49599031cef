"""Host-side preparation of the constraint-level constants the device epilogue needs.

The reference evaluates G(r) / g(r) / S(Q) with numpy float32 expressions inside
``__get_total_Gr`` (Constraints/PairDistributionConstraints.py:847-895),
``__get_total_gr`` (Constraints/PairCorrelationConstraints.py:126-169) and
``__get_total_Sq`` (Constraints/StructureFactorConstraints.py:780-822).  Everything in
those expressions that does not depend on the histogram is a constant of the constraint;
this module evaluates those constants with the same numpy expressions (so they carry the
same float32 roundings and the same numpy scalar-promotion rules as the reference on the
running numpy) and packs them into the C ABI's ``frmc_model_desc``.
"""
import itertools

import numpy as np

from . import _lib as L

FLOAT_TYPE = np.float32
PI = FLOAT_TYPE(np.pi)            # Globals.py:45

KIND_PDF, KIND_PCF, KIND_SQ, KIND_RSQ = 0, 1, 2, 3
KIND_NAMES = {"PDF": KIND_PDF, "PCF": KIND_PCF, "SQ": KIND_SQ, "RSQ": KIND_RSQ}


def elements_pairs(elements):
    """Pair order of the reference's accumulation loop (PairDistributionConstraints.py:485)."""
    return sorted(itertools.combinations_with_replacement(list(elements), 2))


def faber_ziman_weights(n_per_element, element_weights):
    """Normalised pair weights w_ij = c_i c_j b_i b_j / (sum c b)^2 (x2 for i != j), keys "A-B".
    Stands in for pdbparser's get_normalized_weighting (called at
    PairDistributionConstraints.py:487), which is a third-party input to the path."""
    els = list(n_per_element.keys())
    total = float(sum(n_per_element.values()))
    c = {e: n_per_element[e] / total for e in els}
    norm = sum(c[e] * float(element_weights[e]) for e in els) ** 2
    out = {}
    for i, a in enumerate(els):
        for b in els[i:]:
            w = c[a] * c[b] * float(element_weights[a]) * float(element_weights[b]) / norm
            out[a + "-" + b] = FLOAT_TYPE(2.0 * w if a != b else w)
    return out


def gr2sq_matrix(q_values, shell_centers):
    """[hs, nQ] float32 matrix dr*sin(Q r)/Q (StructureFactorConstraints.py:302-312)."""
    Qs = np.asarray(q_values, dtype=FLOAT_TYPE)
    Rs = np.asarray(shell_centers, dtype=FLOAT_TYPE)
    dr = Rs[1] - Rs[0]
    qr = Rs.reshape((-1, 1)) * (np.ones((len(Rs), 1), dtype=FLOAT_TYPE) * Qs)
    return np.ascontiguousarray(dr * (np.sin(qr) / Qs), dtype=FLOAT_TYPE)


def shell_volumes_from_edges(edges):
    """PairDistributionConstraints.py:758 / StructureFactorConstraints.py:350."""
    edges = np.asarray(edges, dtype=FLOAT_TYPE)
    return FLOAT_TYPE(4.0 / 3.) * PI * ((edges[1:]) ** 3 - edges[0:-1] ** 3)


class ModelSpec(object):
    """Everything one constraint contributes to the device epilogue.

    :Parameters:
        #. kind (str): "PDF", "PCF", "SQ" or "RSQ".
        #. elements (list): engine.elements (element index -> name).
        #. n_per_element (dict): engine.numberOfAtomsPerElement.
        #. weighting (dict): the constraint's weightingScheme ("A-B" -> float32).
        #. volume, rho0 (float32): engine.volume, engine.numberDensity.
        #. shell_centers, shell_volumes (float32 arrays, histSize).
        #. experimental (float32 array): experimentalPDF / experimentalSF inside the limits.
        #. data_weights (None, float32 array): usedDataWeights.
        #. shape_array (None, float32 array): shape-function array (PDF/PCF).
        #. scale_factor (float): fitted scale factor.
        #. q_values (None, float32 array): experimental Q values (SQ/RSQ); the Gr2Sq matrix is
           built from them unless ``gr2sq`` is given.
    """

    def __init__(self, kind, elements, n_per_element, weighting, volume, rho0, shell_centers, shell_volumes,
                 experimental, data_weights=None, shape_array=None, scale_factor=1.0, q_values=None, gr2sq=None,
                 sq_exact=True):
        self.kind = KIND_NAMES[kind] if isinstance(kind, str) else int(kind)
        self.elements = list(elements)
        self.n_per_element = dict(n_per_element)
        self.weighting = dict(weighting)
        self.volume = FLOAT_TYPE(volume)
        self.rho0 = FLOAT_TYPE(rho0)
        self.shell_centers = np.ascontiguousarray(shell_centers, dtype=FLOAT_TYPE)
        self.shell_volumes = np.ascontiguousarray(shell_volumes, dtype=FLOAT_TYPE)
        self.experimental = np.ascontiguousarray(experimental, dtype=FLOAT_TYPE)
        self.data_weights = None if data_weights is None else np.ascontiguousarray(data_weights, dtype=FLOAT_TYPE)
        self.shape_array = None if shape_array is None else np.ascontiguousarray(shape_array, dtype=FLOAT_TYPE)
        self.scale_factor = FLOAT_TYPE(scale_factor)
        self.sq_exact = int(sq_exact)
        self.gr2sq = None
        if self.kind in (KIND_SQ, KIND_RSQ):
            if gr2sq is None:
                if q_values is None:
                    raise ValueError("S(Q) models need q_values or a Gr2Sq matrix")
                gr2sq = gr2sq_matrix(q_values, self.shell_centers)
            self.gr2sq = np.ascontiguousarray(gr2sq, dtype=FLOAT_TYPE)
            if self.gr2sq.shape != (self.shell_centers.shape[0], self.experimental.shape[0]):
                raise ValueError("Gr2Sq matrix must be (histSize, nQ)")

    def pair_table(self):
        """(idi, idj, wij, Dij) per pair in the reference's loop order."""
        pa, pb, pw, pD = [], [], [], []
        for pair in elements_pairs(self.elements):
            wij = self.weighting.get(pair[0] + "-" + pair[1], None)
            if wij is None:
                wij = self.weighting[pair[1] + "-" + pair[0]]
            ni = self.n_per_element[pair[0]]
            nj = self.n_per_element[pair[1]]
            idi = self.elements.index(pair[0])
            idj = self.elements.index(pair[1])
            if idi == idj:
                Nij = ni * (ni - 1) / 2.0                       # PairDistributionConstraints.py:867
            else:
                Nij = ni * nj                                   # :872
            Dij = FLOAT_TYPE(Nij / self.volume)                 # :868 (PCF/SQ omit the cast: same value, numpy>=2)
            pa.append(idi); pb.append(idj); pw.append(FLOAT_TYPE(wij)); pD.append(Dij)
        return (np.array(pa, dtype=np.int32), np.array(pb, dtype=np.int32),
                np.array(pw, dtype=FLOAT_TYPE), np.array(pD, dtype=FLOAT_TYPE))

    def prefactor(self):
        """(4.*PI*shellCenters*rho0) as the reference writes it (PDF :881, SQ :808)."""
        if self.kind in (KIND_SQ, KIND_RSQ):
            return np.ascontiguousarray(FLOAT_TYPE(4.) * PI * self.shell_centers * self.rho0, dtype=FLOAT_TYPE)
        return np.ascontiguousarray(4. * PI * self.shell_centers * self.rho0, dtype=FLOAT_TYPE)

    def build_desc(self):
        """Returns (ModelDesc, keepalive) -- keepalive holds the numpy arrays the desc points into."""
        pa, pb, pw, pD = self.pair_table()
        pref = self.prefactor()
        keep = [pa, pb, pw, pD, pref, self.shell_volumes, self.experimental, self.data_weights, self.shape_array,
                self.gr2sq]
        d = L.ModelDesc()
        d.kind = self.kind
        d.n_pairs = pa.shape[0]
        d.pair_a = L.ptr(pa, L.c_i32p)
        d.pair_b = L.ptr(pb, L.c_i32p)
        d.pair_w = L.ptr(pw, L.c_f32p)
        d.pair_D = L.ptr(pD, L.c_f32p)
        d.shell_volumes = L.ptr(self.shell_volumes, L.c_f32p)
        d.prefactor = L.ptr(pref, L.c_f32p)
        d.shape = L.ptr(self.shape_array, L.c_f32p)
        d.scale = float(self.scale_factor)
        d.n_out = self.experimental.shape[0]
        d.experimental = L.ptr(self.experimental, L.c_f32p)
        d.data_weights = L.ptr(self.data_weights, L.c_f32p)
        d.gr2sq = L.ptr(self.gr2sq, L.c_f32p)
        d.sq_exact = int(self.sq_exact)
        return d, keep
