"""Coordination-number constraint on the CUDA counting kernels.

What the reference's ``AtomicCoordinationNumberConstraint`` (Constraints/AtomicCoordinationConstraints.py) keeps and
computes, expressed for this backend:

* ``data`` -- float32, one entry per shell definition: the number of (core, shell) neighbour pairs of the whole system.
  The whole-system count meets every pair from both ends, hence the halving (:491-505).
* a move changes ``data`` by the counts of the moved atoms after the move minus their counts before it (:519-577).
* the standard error penalises the MEAN coordination number ``data / number of core atoms`` of a definition for lying
  outside ``[minAtoms, maxAtoms]``, linearly and weighted, summed over the definitions in order in float32 (:404-461).

The engine-facing protocol (``compute_data`` / ``compute_before_move`` / ``compute_after_move`` / ``accept_move`` /
``reject_move``; Core/Constraint.py:732-748) is kept, with one staged move at a time.  The counts come from
``fullrmc_b200.Core.atomic_coordination`` (one launch per call over a flat task table, bit-identical to the reference's
Cython loops).  Two ways to run:

* stateless (``store=None``): coordinates and lists travel with every call, like the reference's own functions;
* on the device store (``store=DeviceStore``): the definitions are registered once on the store whose atoms the histogram
  constraints move, and the before / after counts of a move come from ONE launch over the resident records
  (``DeviceStore.coordination_move``, csrc/storecoord.cu) -- no coordinate or list upload; the store owns the
  coordinates (``boxCoordinates`` is only read by ``compute_data`` when no store is given).
"""
import collections

import numpy as np

FLOAT_TYPE = np.float32
INT_TYPE = np.int32

_StagedMove = collections.namedtuple("_StagedMove", "indexes before after data error")


def _membership(lists, n_atoms):
    """per atom the definitions whose list names it, in definition order (what the counting kernels' task builder wants)"""
    out = [[] for _ in range(n_atoms)]
    for d, atoms in enumerate(lists):
        for a in atoms:
            out[a].append(d)
    return out


class DeviceAtomicCoordinationNumberConstraint(object):
    """:Parameters:
        #. boxCoordinates, basisVectors, isPBC: the engine arrays.  boxCoordinates is read at every call and never
           written: the after-move counts are taken on a patched copy.
        #. coresIndexes, shellsIndexes, lowerShells, upperShells, minAtoms, maxAtoms, weights: per definition the sorted
           int32 core and shell atom lists, the shell bounds, the allowed range of the mean coordination number and
           the weight -- what set_coordination_number_definition derives (:242-420).
    """

    def __init__(self, boxCoordinates, basisVectors, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells, minAtoms,
                 maxAtoms, weights=None, kernels=None, store=None):
        self._store = store
        if kernels is None:
            from .Core import atomic_coordination as kernels
        self._count = kernels
        self.boxCoordinates = boxCoordinates
        self.basisVectors = np.ascontiguousarray(basisVectors, dtype=FLOAT_TYPE)
        self.isPBC = bool(isPBC)
        as_lists = lambda seq: [np.ascontiguousarray(x, dtype=INT_TYPE) for x in seq]
        as_f32 = lambda seq: np.array([FLOAT_TYPE(x) for x in seq], dtype=FLOAT_TYPE)
        self.coresIndexes, self.shellsIndexes = as_lists(coresIndexes), as_lists(shellsIndexes)
        n_def = len(self.coresIndexes)
        self.lowerShells, self.upperShells = list(as_f32(lowerShells)), list(as_f32(upperShells))
        self.minAtoms, self.maxAtoms = as_f32(minAtoms), as_f32(maxAtoms)
        self.weights = np.ones(n_def, FLOAT_TYPE) if weights is None else as_f32(weights)
        sizes = {len(self.shellsIndexes), len(self.lowerShells), len(self.upperShells), self.minAtoms.shape[0],
                 self.maxAtoms.shape[0], self.weights.shape[0]}
        if sizes != {n_def}:
            raise ValueError("every per-definition argument needs one entry per shell definition (%d)" % n_def)
        n_atoms = boxCoordinates.shape[0]
        self.asCoreDefIdxs = _membership(self.coresIndexes, n_atoms)
        self.inShellDefIdxs = _membership(self.shellsIndexes, n_atoms)
        self.numberOfCores = np.array([c.shape[0] for c in self.coresIndexes], dtype=FLOAT_TYPE)
        self.data = None
        self.standardError = None
        self._staged = None
        self.tried = 0
        self.accepted = 0
        self._sc = None
        if store is not None:
            self._sc = store.coordination_add(self.coresIndexes, self.shellsIndexes, self.lowerShells, self.upperShells)

    # ---------------------------------------------------------------- counting
    def _definitions(self):
        return dict(basis=self.basisVectors, isPBC=self.isPBC, coresIndexes=self.coresIndexes, shellsIndexes=self.shellsIndexes,
                    lowerShells=self.lowerShells, upperShells=self.upperShells, asCoreDefIdxs=self.asCoreDefIdxs,
                    inShellDefIdxs=self.inShellDefIdxs, ncores=1)

    def _counts_of(self, indexes, coordinates):
        """neighbour pairs the listed atoms take part in, per definition"""
        out = np.zeros(len(self.coresIndexes), dtype=FLOAT_TYPE)
        self._count.multi_atoms_coord_number_coords(indexes=indexes, boxCoords=coordinates, coordNumData=out, **self._definitions())
        return out

    def compute_standard_error(self, data):
        """weighted linear penalty of the mean coordination numbers outside their ranges; terms added in definition order
        in float32 (a definition inside its range adds an exact zero), which is the order of the reference's loop"""
        mean = (np.asarray(data, dtype=FLOAT_TYPE) / self.numberOfCores).astype(FLOAT_TYPE)
        below, above = mean < self.minAtoms, mean > self.maxAtoms
        terms = np.zeros(mean.shape[0], dtype=FLOAT_TYPE)
        terms[below] = (self.weights * (self.minAtoms - mean))[below]
        only_above = above & ~below
        terms[only_above] = (self.weights * (mean - self.maxAtoms))[only_above]
        return FLOAT_TYPE(np.add.accumulate(terms, dtype=FLOAT_TYPE)[-1]) if terms.shape[0] else FLOAT_TYPE(0.0)

    # ---------------------------------------------------------------- the engine-facing protocol
    def compute_data(self, update=True):
        pairs = np.zeros(len(self.coresIndexes), dtype=FLOAT_TYPE)
        coords = self.boxCoordinates if self._store is None else self._store.get_coords()
        self._count.all_atoms_coord_number_coords(boxCoords=coords, coordNumData=pairs, **self._definitions())
        pairs /= FLOAT_TYPE(2.)                              # every neighbour pair was met from both of its atoms
        error = self.compute_standard_error(pairs)
        if update:
            self.data, self.standardError, self._staged = pairs, error, None
        return pairs, error

    def compute_before_move(self, realIndexes, relativeIndexes):
        idx = np.ascontiguousarray(relativeIndexes, dtype=INT_TYPE)
        if self._store is not None:
            # before and after come from one launch over the store once the moved coordinates are known
            self._staged = _StagedMove(idx, None, None, None, None)
            return
        self._staged = _StagedMove(idx, self._counts_of(idx, self.boxCoordinates), None, None, None)

    def compute_after_move(self, realIndexes, relativeIndexes, movedBoxCoordinates):
        idx = np.ascontiguousarray(relativeIndexes, dtype=INT_TYPE)
        if self._store is not None:
            counts = self._store.coordination_move(self._sc, idx, movedBoxCoordinates).astype(FLOAT_TYPE)
            data = self.data - counts[0] + counts[1]
            self._staged = _StagedMove(idx, counts[0], counts[1], data, self.compute_standard_error(data))
            self._moved = (idx, np.ascontiguousarray(movedBoxCoordinates, dtype=FLOAT_TYPE))
            self.tried += 1
            return
        if self._staged is None or not np.array_equal(self._staged.indexes, idx):
            self.compute_before_move(realIndexes, relativeIndexes)
        patched = np.array(self.boxCoordinates, dtype=FLOAT_TYPE)
        patched[idx] = movedBoxCoordinates
        after = self._counts_of(idx, patched)
        data = self.data - self._staged.before + after
        self._staged = self._staged._replace(after=after, data=data, error=self.compute_standard_error(data))
        self.tried += 1

    def accept_move(self, realIndexes, relativeIndexes):
        self.data, self.standardError = self._staged.data, self._staged.error
        self._staged = None
        self.accepted += 1
        if self._store is not None and self._store.n_models == 0:
            self._store.move_atoms(*self._moved)                  # nobody else commits the move on a store without models

    def reject_move(self, realIndexes, relativeIndexes):
        self._staged = None

    # ---------------------------------------------------------------- the reference's attribute names for the staged move
    @property
    def afterMoveStandardError(self):
        return None if self._staged is None else self._staged.error

    @property
    def activeAtomsDataBeforeMove(self):
        return None if self._staged is None else self._staged.before

    @property
    def activeAtomsDataAfterMove(self):
        return None if self._staged is None else self._staged.after
