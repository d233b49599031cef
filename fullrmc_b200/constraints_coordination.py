"""Device-backed mirror of ``AtomicCoordinationNumberConstraint``'s five methods.

Reference class: Constraints/AtomicCoordinationConstraints.py.  State, method names and arithmetic follow it:

* data = float32 [number of definitions]: per definition the number of (core, shell) neighbour pairs inside the shell,
  i.e. the whole-system count of ``all_atoms_coord_number_coords`` halved, because every pair is met from both of its
  ends (:491-505);
* a move's contribution is ``multi_atoms_coord_number_coords`` of the moved atoms before and after (:519-577), and the
  data after the move is ``data - before + after`` (:575);
* standardError = sum over the definitions of ``weight * (min - CN)`` or ``weight * (CN - max)`` where the mean
  coordination number ``CN = data / number of core atoms`` lies outside ``[min, max]`` (:404-461);
* accept_move / reject_move commit or drop the staged data (:580-612).

The counting functions are ``fullrmc_b200.Core.atomic_coordination`` (CUDA, bit-identical to the reference's Cython);
this first version is stateless -- the coordinates travel with every call -- like the reference's own functions.
"""
import numpy as np

FLOAT_TYPE = np.float32
INT_TYPE = np.int32


class DeviceAtomicCoordinationNumberConstraint(object):
    """:Parameters:
        #. boxCoordinates, basisVectors, isPBC: the engine arrays (boxCoordinates is read at every call; compute_after_move
           writes the moved coordinates into it and restores them, as the reference does, :555-572).
        #. coresIndexes, shellsIndexes, lowerShells, upperShells, minAtoms, maxAtoms, weights: what the reference derives
           in set_coordination_number_definition (:242-420): per definition the sorted int32 core and shell atom lists,
           the shell bounds, the allowed range of the mean coordination number and the weight.
    """

    def __init__(self, boxCoordinates, basisVectors, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells, minAtoms,
                 maxAtoms, weights=None, kernels=None):
        if kernels is None:
            from .Core import atomic_coordination as kernels
        self._kernels = kernels
        self.boxCoordinates = boxCoordinates
        self.basisVectors = np.ascontiguousarray(basisVectors, dtype=FLOAT_TYPE)
        self.isPBC = bool(isPBC)
        self.coresIndexes = [np.ascontiguousarray(c, dtype=INT_TYPE) for c in coresIndexes]
        self.shellsIndexes = [np.ascontiguousarray(s, dtype=INT_TYPE) for s in shellsIndexes]
        ndef = len(self.coresIndexes)
        self.lowerShells = [FLOAT_TYPE(x) for x in lowerShells]
        self.upperShells = [FLOAT_TYPE(x) for x in upperShells]
        self.minAtoms = [FLOAT_TYPE(x) for x in minAtoms]
        self.maxAtoms = [FLOAT_TYPE(x) for x in maxAtoms]
        self.weights = np.ones(ndef, FLOAT_TYPE) if weights is None else np.array(weights, dtype=FLOAT_TYPE)
        assert len(self.shellsIndexes) == ndef and len(self.lowerShells) == ndef and len(self.upperShells) == ndef
        assert len(self.minAtoms) == ndef and len(self.maxAtoms) == ndef and self.weights.shape == (ndef,)
        # per atom: the definitions it is a core of / in the shell of (:385-400)
        n = boxCoordinates.shape[0]
        self.asCoreDefIdxs = [[] for _ in range(n)]
        self.inShellDefIdxs = [[] for _ in range(n)]
        for defIdx in range(ndef):
            for atIdx in self.coresIndexes[defIdx]:
                self.asCoreDefIdxs[atIdx].append(defIdx)
            for atIdx in self.shellsIndexes[defIdx]:
                self.inShellDefIdxs[atIdx].append(defIdx)
        self.numberOfCores = np.array([len(c) for c in self.coresIndexes], dtype=FLOAT_TYPE)
        self.data = None
        self.standardError = None
        self.afterMoveStandardError = None
        self.activeAtomsDataBeforeMove = None
        self.activeAtomsDataAfterMove = None
        self._dataAfterMove = None
        self.tried = 0
        self.accepted = 0

    def _lists(self):
        return dict(basis=self.basisVectors, isPBC=self.isPBC, coresIndexes=self.coresIndexes, shellsIndexes=self.shellsIndexes,
                    lowerShells=self.lowerShells, upperShells=self.upperShells, asCoreDefIdxs=self.asCoreDefIdxs,
                    inShellDefIdxs=self.inShellDefIdxs, ncores=1)

    def compute_standard_error(self, data):
        """:404-461, term by term (numpy float32 scalars, so the sum runs in float32 as the reference's does)"""
        coordNum = data / self.numberOfCores
        StdErr = 0.
        for idx, cn in enumerate(coordNum):
            if cn < self.minAtoms[idx]:
                StdErr += self.weights[idx] * (self.minAtoms[idx] - cn)
            elif cn > self.maxAtoms[idx]:
                StdErr += self.weights[idx] * (cn - self.maxAtoms[idx])
        return StdErr

    def compute_data(self, update=True):
        """:473-517"""
        coordNumData = np.zeros(len(self.coresIndexes), dtype=FLOAT_TYPE)
        self._kernels.all_atoms_coord_number_coords(boxCoords=self.boxCoordinates, coordNumData=coordNumData, **self._lists())
        coordNumData /= FLOAT_TYPE(2.)
        stdError = self.compute_standard_error(data=coordNumData)
        if update:
            self.data = coordNumData
            self.activeAtomsDataBeforeMove = None
            self.activeAtomsDataAfterMove = None
            self.standardError = stdError
        return coordNumData, stdError

    def compute_before_move(self, realIndexes, relativeIndexes):
        """:519-543"""
        beforeMoveData = np.zeros(self.data.shape, dtype=self.data.dtype)
        self._kernels.multi_atoms_coord_number_coords(indexes=np.ascontiguousarray(relativeIndexes, dtype=INT_TYPE),
                                                      boxCoords=self.boxCoordinates, coordNumData=beforeMoveData, **self._lists())
        self.activeAtomsDataBeforeMove = beforeMoveData
        self.activeAtomsDataAfterMove = None

    def compute_after_move(self, realIndexes, relativeIndexes, movedBoxCoordinates):
        """:545-578"""
        boxData = np.array(self.boxCoordinates[relativeIndexes], dtype=FLOAT_TYPE)
        self.boxCoordinates[relativeIndexes] = movedBoxCoordinates
        afterMoveData = np.zeros(self.data.shape, dtype=self.data.dtype)
        try:
            self._kernels.multi_atoms_coord_number_coords(indexes=np.ascontiguousarray(relativeIndexes, dtype=INT_TYPE),
                                                          boxCoords=self.boxCoordinates, coordNumData=afterMoveData, **self._lists())
        finally:
            self.boxCoordinates[relativeIndexes] = boxData
        self.activeAtomsDataAfterMove = afterMoveData
        self._dataAfterMove = self.data - self.activeAtomsDataBeforeMove + self.activeAtomsDataAfterMove
        self.afterMoveStandardError = self.compute_standard_error(data=self._dataAfterMove)
        self.tried += 1

    def accept_move(self, realIndexes, relativeIndexes):
        """:580-598"""
        self.data = self._dataAfterMove
        self.activeAtomsDataBeforeMove = None
        self.activeAtomsDataAfterMove = None
        self.standardError = self.afterMoveStandardError
        self.afterMoveStandardError = None
        self.accepted += 1

    def reject_move(self, realIndexes, relativeIndexes):
        """:600-612"""
        self.activeAtomsDataBeforeMove = None
        self.activeAtomsDataAfterMove = None
        self.afterMoveStandardError = None
