"""Device-backed mirror of the molecular distance constraints' five methods.

Reference classes: ``InterMolecularDistanceConstraint`` / ``IntraMolecularDistanceConstraint``
(Constraints/DistanceConstraints.py:966-1203) on top of ``_MolecularDistanceConstraint`` (:470-820) and
``_DistanceConstraint`` (:23-468).  State, method names and arithmetic follow the reference:

* data = {"number": int32 [nT,nT,1], "distanceSum": float32 [nT,nT,1]} of the inter- (or intra-) molecular pairs
  closer than their type pair's limit, distances reduced to ``|upper - d|`` (the fixed flags of :483-486);
* a move's contribution is ``multiple(all atoms) - full(the group alone)`` before and after (:606-737);
* standardError = float32(np.sum(distanceSum_sym / number_sym)) over the type pairs (:385-428);
* a rigid (non-flexible) constraint rejects a step when any cell's count grows (:361-383).

The two kernels are ``fullrmc_b200.Core.atomic_distances`` (CUDA, bit-identical to the reference's Cython, sums
included).  Two modes:

* stateless (``store=None``): the coordinates travel with every call, like the reference's own functions;
* on the device store (``store=DeviceStore``): the constraint is registered once on the store
  whose atoms the histogram constraints move, and a move's four quantities (M and F, before and after) come from ONE
  pass over the resident records (``DeviceStore.distance_move``, csrc/storedist.cu) -- no coordinate upload; the
  engine's boxCoordinates array is not read at all in the Monte-Carlo loop.
"""
import numpy as np

FLOAT_TYPE = np.float32
INT_TYPE = np.int32


class DeviceMolecularDistanceConstraint(object):
    """:Parameters:
        #. boxCoordinates, basisVectors, isPBC, moleculesIndex: the engine arrays (boxCoordinates is read at every
           call and never written: the engine owns it, Engine.py:3337-3338).
        #. typesIndex, numberOfTypes, lowerLimitArray, upperLimitArray, typePairsIndex: what the reference constraint
           derives in set_type_definition / set_pairs_distance (DistanceConstraints.py:195-359).
        #. interMolecular (bool): True for the Inter class, False for the Intra class.
        #. flexible (bool): the reference's flag of the same name.
    """

    def __init__(self, boxCoordinates, basisVectors, isPBC, moleculesIndex, typesIndex, numberOfTypes, lowerLimitArray,
                 upperLimitArray, typePairsIndex, interMolecular=True, flexible=True, kernels=None, store=None):
        if kernels is None:
            from .Core import atomic_distances as kernels
        self._kernels = kernels
        self._store = store
        self._pending = None                 # store mode: the group of the move being evaluated
        self.boxCoordinates = boxCoordinates
        self.basisVectors = np.ascontiguousarray(basisVectors, dtype=FLOAT_TYPE)
        self.isPBC = bool(isPBC)
        self.moleculesIndex = np.ascontiguousarray(moleculesIndex, dtype=INT_TYPE)
        self.typesIndex = np.ascontiguousarray(typesIndex, dtype=INT_TYPE)
        self.numberOfTypes = int(numberOfTypes)
        self.lowerLimitArray = np.ascontiguousarray(lowerLimitArray, dtype=FLOAT_TYPE)
        self.upperLimitArray = np.ascontiguousarray(upperLimitArray, dtype=FLOAT_TYPE)
        self.typePairsIndex = np.ascontiguousarray(typePairsIndex, dtype=INT_TYPE)
        self._interMolecular = bool(interMolecular)
        self._intraMolecular = not self._interMolecular
        self.flexible = bool(flexible)
        # the fixed flags of _MolecularDistanceConstraint.__init__ (:483-486)
        self._flags = dict(interMolecular=self._interMolecular, intraMolecular=self._intraMolecular, reduceDistance=False,
                           reduceDistanceToUpper=True, reduceDistanceToLower=False, countWithinLimits=True)
        self.data = None
        self.standardError = None
        self.afterMoveStandardError = None
        self.activeAtomsDataBeforeMove = None
        self.activeAtomsDataAfterMove = None
        self.tried = 0
        self.accepted = 0
        if store is not None:
            self._sd = store.distance_add(self.typesIndex, self.numberOfTypes, self.lowerLimitArray, self.upperLimitArray,
                                          **{k: v for k, v in self._flags.items()})

    # ---------------------------------------------------------------- helpers
    def _pick(self, result):
        nintra, dintra, ninter, dinter = result
        return (ninter, dinter) if self._interMolecular else (nintra, dintra)

    def _system(self, coords):
        return dict(boxCoords=coords, basis=self.basisVectors, isPBC=self.isPBC, numberOfElements=self.numberOfTypes,
                    lowerLimit=self.lowerLimitArray, upperLimit=self.upperLimitArray)

    def _move_contribution(self, coords, relativeIndexes):
        """multiple(all N) - full(group alone) (:606-662)"""
        idx = np.ascontiguousarray(relativeIndexes, dtype=INT_TYPE)
        numberM, sumM = self._pick(self._kernels.multiple_atomic_distances_coords(
            indexes=idx, moleculeIndex=self.moleculesIndex, elementIndex=self.typesIndex, allAtoms=True,
            **self._system(coords), **self._flags))
        numberF, sumF = self._pick(self._kernels.full_atomic_distances_coords(
            moleculeIndex=np.ascontiguousarray(self.moleculesIndex[idx]), elementIndex=np.ascontiguousarray(self.typesIndex[idx]),
            **self._system(np.ascontiguousarray(coords[idx])), **self._flags))
        return {"number": numberM - numberF, "distanceSum": sumM - sumF}

    def _get_constraint_value(self, data=None):
        """per type pair the mean reduced distance of the counted pairs: both orderings of the pair summed, then
        sum / count where anything was counted (:415-428)"""
        data = self.data if data is None else data
        first, second = self.typePairsIndex[:, 0], self.typePairsIndex[:, 1]
        both = lambda a: (a[first, second] + a[second, first]).reshape(-1)
        counts, sums = both(data["number"]), both(data["distanceSum"])
        return np.divide(sums, counts, out=sums.copy(), where=counts != 0)

    def _compute_standard_error(self, distances):
        return FLOAT_TYPE(np.sum(distances))

    # ---------------------------------------------------------------- the five methods
    def _coords(self):
        if self._store is not None:
            return self._store.get_coords()                       # the store owns the coordinates
        return np.ascontiguousarray(self.boxCoordinates, dtype=FLOAT_TYPE)

    def compute_data(self, update=True):
        coords = self._coords()
        number, distanceSum = self._pick(self._kernels.full_atomic_distances_coords(
            moleculeIndex=self.moleculesIndex, elementIndex=self.typesIndex, **self._system(coords), **self._flags))
        data = {"number": number, "distanceSum": distanceSum}
        stdError = self._compute_standard_error(self._get_constraint_value(data))
        if update:
            self.data = data
            self.activeAtomsDataBeforeMove = self.activeAtomsDataAfterMove = None
            self.standardError = stdError
        return data, stdError

    def compute_before_move(self, realIndexes, relativeIndexes):
        if self._store is not None:
            # before and after come from one pass over the store once the moved coordinates are known
            self._pending = np.ascontiguousarray(relativeIndexes, dtype=INT_TYPE)
            self.activeAtomsDataBeforeMove = self.activeAtomsDataAfterMove = None
            return
        coords = np.ascontiguousarray(self.boxCoordinates, dtype=FLOAT_TYPE)
        self.activeAtomsDataBeforeMove = self._move_contribution(coords, relativeIndexes)
        self.activeAtomsDataAfterMove = None

    def compute_after_move(self, realIndexes, relativeIndexes, movedBoxCoordinates):
        if self._store is not None:
            idx = np.ascontiguousarray(relativeIndexes, dtype=INT_TYPE)
            counts, sums = self._store.distance_move(self._sd, idx, movedBoxCoordinates)
            part = 1 if self._interMolecular else 0
            # multiple(all N) - full(group alone), before and after (:606-737)
            self.activeAtomsDataBeforeMove = {"number": counts[0, part] - counts[1, part], "distanceSum": sums[0, part] - sums[1, part]}
            self.activeAtomsDataAfterMove = {"number": counts[2, part] - counts[3, part], "distanceSum": sums[2, part] - sums[3, part]}
            self._moved = (idx, np.ascontiguousarray(movedBoxCoordinates, dtype=FLOAT_TYPE))
        else:
            coords = np.array(self.boxCoordinates, dtype=FLOAT_TYPE)                   # a copy: the engine's array stays as it is
            coords[relativeIndexes] = movedBoxCoordinates
            self.activeAtomsDataAfterMove = self._move_contribution(coords, relativeIndexes)
        number = self.data["number"] - self.activeAtomsDataBeforeMove["number"] + self.activeAtomsDataAfterMove["number"]
        distanceSum = self.data["distanceSum"] - self.activeAtomsDataBeforeMove["distanceSum"] + self.activeAtomsDataAfterMove["distanceSum"]
        self.afterMoveStandardError = self._compute_standard_error(self._get_constraint_value({"number": number, "distanceSum": distanceSum}))
        self.tried += 1

    def should_step_get_rejected(self, standardError):
        """:361-383; the flexible variant is RigidConstraint's rule (Core/Constraint.py): reject when the error grows"""
        if self.flexible:
            return bool(standardError > self.standardError)
        return bool(np.any(self.activeAtomsDataAfterMove["number"] > self.activeAtomsDataBeforeMove["number"]))

    def accept_move(self, realIndexes, relativeIndexes):
        number = self.data["number"] - self.activeAtomsDataBeforeMove["number"] + self.activeAtomsDataAfterMove["number"]
        distanceSum = self.data["distanceSum"] - self.activeAtomsDataBeforeMove["distanceSum"] + self.activeAtomsDataAfterMove["distanceSum"]
        self.data = {"number": number, "distanceSum": distanceSum}
        self.activeAtomsDataBeforeMove = self.activeAtomsDataAfterMove = None
        self.standardError = self.afterMoveStandardError
        self.afterMoveStandardError = None
        self.accepted += 1
        if self._store is not None and self._store.n_models == 0:
            self._store.move_atoms(*self._moved)                  # nobody else commits the move on a store without models

    def reject_move(self, realIndexes, relativeIndexes):
        self.activeAtomsDataBeforeMove = self.activeAtomsDataAfterMove = None
        self.afterMoveStandardError = None
