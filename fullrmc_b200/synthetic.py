"""Seeded synthetic systems for the BASELINE.json configs 4 and 5.

No data files: everything is generated from ``numpy.random.default_rng(seed)``.
"""
import numpy as np

from .model import FLOAT_TYPE, PI, faber_ziman_weights, shell_volumes_from_edges

ELEMENTS = ["O", "Si", "Ti", "Ni", "Zr"]          # Z = 8, 14, 22, 28, 40
ATOMIC_NUMBERS = {"O": 8, "Si": 14, "Ti": 22, "Ni": 28, "Zr": 40}

CFG4_BASIS = np.array([[100, 0, 0], [15, 98, 0], [-10, 20, 95]], dtype=np.float32)
CFG5_EDGE = 215.44


class System(object):
    """A synthetic periodic (or infinite) system with the arrays the extension functions take."""

    def __init__(self, box, basis, isPBC, mol, el, n_elements):
        self.boxCoords = box
        self.basis = basis
        self.isPBC = isPBC
        self.moleculeIndex = mol
        self.elementIndex = el
        self.numberOfElements = n_elements
        self.numberOfAtoms = box.shape[0]
        self.elements = ELEMENTS[:n_elements]
        counts = np.bincount(el, minlength=n_elements)
        self.numberOfAtomsPerElement = {self.elements[i]: int(counts[i]) for i in range(n_elements)}
        self.volume = FLOAT_TYPE(abs(np.linalg.det(basis.astype(np.float64)))) if isPBC else FLOAT_TYPE(box.shape[0] / 0.0333679)
        self.numberDensity = FLOAT_TYPE(box.shape[0]) / FLOAT_TYPE(self.volume)
        weights = {e: FLOAT_TYPE(ATOMIC_NUMBERS[e]) for e in self.elements}
        self.weighting = faber_ziman_weights(self.numberOfAtomsPerElement, weights)

    def hist_kwargs(self):
        return dict(basis=self.basis, isPBC=self.isPBC, moleculeIndex=self.moleculeIndex,
                    elementIndex=self.elementIndex, numberOfElements=self.numberOfElements)


def random_system(n, seed, basis, n_elements=5, molecule_size=1, isPBC=True, spread=None):
    """Uniform fractional coordinates in [0,1) (or [-spread, 1+spread) when given)."""
    rng = np.random.default_rng(seed)
    box = rng.random((n, 3), dtype=np.float32)
    if spread:
        box = (box * np.float32(1 + 2 * spread) - np.float32(spread)).astype(np.float32)
    if not isPBC:
        box = (box.astype(np.float64) @ basis.astype(np.float64)).astype(np.float32)
    el = rng.integers(0, n_elements, n).astype(np.int32)
    mol = (np.arange(n, dtype=np.int64) // molecule_size).astype(np.int32)
    return System(box, np.ascontiguousarray(basis, dtype=np.float32), isPBC, mol, el, n_elements)


def cfg4(n=100000, seed=4):
    """synthetic 100k-atom 5-element triclinic box (BASELINE.json configs[3])."""
    return random_system(n, seed, CFG4_BASIS)


def cfg5(n=1000000, seed=5):
    """synthetic 1M-atom cubic box (BASELINE.json configs[4]); the edge scales with n^(1/3)
    so the density stays 0.1 atoms/A^3 when a smaller n is requested."""
    edge = CFG5_EDGE * (n / 1.0e6) ** (1.0 / 3.0)
    basis = np.diag([edge, edge, edge]).astype(np.float32)
    return random_system(n, seed, basis)


class RGrid(object):
    """r-grid of the synthetic configs: rmin=0, bin=0.02, hs=1000 (rmax=20)."""

    def __init__(self, rmin=0.0, bin=0.02, hs=1000):
        self.bin = FLOAT_TYPE(bin)
        self.hs = int(hs)
        self.edges = (FLOAT_TYPE(rmin) + self.bin * np.arange(hs + 1, dtype=np.float64)).astype(FLOAT_TYPE)
        self.minDistance = FLOAT_TYPE(self.edges[0])
        self.maxDistance = FLOAT_TYPE(self.edges[-1])
        self.shellCenters = ((self.edges[0:-1] + self.edges[1:]) / FLOAT_TYPE(2.)).astype(FLOAT_TYPE)
        self.shellVolumes = shell_volumes_from_edges(self.edges)

    def kwargs(self):
        return dict(minDistance=self.minDistance, maxDistance=self.maxDistance, bin=self.bin, histSize=self.hs)


def q_values(qmin=0.5, qmax=20.0, nq=400):
    return np.linspace(qmin, qmax, nq).astype(FLOAT_TYPE)


def smooth_target(n, seed, center):
    """experimental stand-in for the synthetic configs: smooth noise around the ideal-gas value
    (0 for G(r), 1 for S(Q)), so that Metropolis acceptance is mixed rather than degenerate"""
    rng = np.random.default_rng(seed)
    return (center + 0.02 * np.convolve(rng.standard_normal(n + 20), np.ones(21) / 21.0, "valid")).astype(np.float32)


def translation_proposals(system, n_moves, seed, sigma=0.1):
    """Single-atom Gaussian translations (sigma in Angstrom) expressed in box coordinates:
    (atom index, movedBox[1,3]) pairs, generated up front so GPU and CPU arms see the same moves.
    The moved coordinates are relative to the START configuration of each atom; callers that
    accept moves must re-base (see bench.py)."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, system.numberOfAtoms, n_moves).astype(np.int32)
    inv = np.linalg.inv(system.basis.astype(np.float64))
    disp = rng.normal(0.0, sigma, (n_moves, 3)) @ inv
    return idx, disp.astype(np.float32)
