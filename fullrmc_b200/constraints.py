"""Device-backed mirrors of the three hot-path constraints' five methods.

Reference classes (method names, argument meaning and state attributes are kept):

* PairDistributionConstraint   Constraints/PairDistributionConstraints.py:1001-1166
* PairCorrelationConstraint    Constraints/PairCorrelationConstraints.py:263-392
* StructureFactorConstraint / ReducedStructureFactorConstraint
                               Constraints/StructureFactorConstraints.py:933-1096, :1235-1260

``compute_data / compute_before_move / compute_after_move / accept_move / reject_move`` are
what ``Engine.__on_runtime_step_try_move`` calls (Engine.py:3302-3338).  Here they drive ONE
shared :class:`~fullrmc_b200.store.DeviceStore`: every constraint registered on a
:class:`DeviceBackend` is evaluated by the same pass over the device-resident coordinates
(SURVEY.md section 7.2-6b), so a move costs one kernel launch however many constraints
listen.  The Engine keeps calling the constraints one after another; the first
``compute_after_move`` of a move triggers the device evaluation, the others read its result,
and ``accept_move`` / ``reject_move`` resolve the staged proposal once.

These classes stand alone (the reference's own classes need pdbparser / pyrep, which the
GPU box does not have); INTEGRATION.md shows the three-line subclass a fullrmc maintainer
writes to put them behind the real constraint classes.
"""
import numpy as np

from .model import FLOAT_TYPE, PI, ModelSpec, shell_volumes_from_edges, faber_ziman_weights
from .store import DeviceStore


class DeviceBackend(object):
    """One device store per engine; constraints register their r-grid and model on it."""

    def __init__(self, boxCoordinates, basisVectors, isPBC, moleculesIndex, elementsIndex, elements,
                 numberOfAtomsPerElement, volume, numberDensity, device=None, persistent=False):
        self.elements = list(elements)
        self.numberOfAtomsPerElement = dict(numberOfAtomsPerElement)
        self.volume = FLOAT_TYPE(volume)
        self.numberDensity = FLOAT_TYPE(numberDensity)
        self.store = DeviceStore(boxCoordinates, basisVectors, isPBC, moleculesIndex, elementsIndex,
                                 len(self.elements), device=device)
        if persistent:
            # one resident kernel serves the whole run of moves (DeviceStore.set_persistent); same results
            self.store.set_persistent(True)
        self.constraints = []
        self._grids = {}
        self.basisVectors = np.ascontiguousarray(basisVectors, dtype=FLOAT_TYPE)
        self.isPBC = bool(isPBC)
        self.moleculesIndex = np.ascontiguousarray(moleculesIndex, dtype=np.int32)
        self.elementsIndex = np.ascontiguousarray(elementsIndex, dtype=np.int32)
        self.accepted = 0            # the engine's accepted-move count (refit and shape-refresh schedules)
        self._move = None            # (indexes, moved) of the proposal being evaluated
        self._move_src = None        # the argument objects themselves (identity shortcut within one step)
        self._chi2 = None            # chi^2 per model of the staged proposal
        self._resolved = True
        self._dirty = True           # committed data need a compute_data pass
        self.allowFittingScaleFactor = False   # engine._RT_moveGenerator.allowFittingScaleFactor of a remove generator
        self._amputation = None      # (relative index, chi^2 per model) of the removal being tried

    # -------------------------------------------------------------- registration
    def _grid(self, minDistance, maxDistance, bin, histSize):
        key = (float(minDistance), float(maxDistance), float(bin), int(histSize))
        if key not in self._grids:
            self._grids[key] = self.store.add_grid(minDistance, maxDistance, bin, histSize)
        return self._grids[key]

    def _register(self, constraint, grid_key, spec):
        g = self._grid(*grid_key)
        constraint._model = self.store.add_model(g, spec)
        constraint._grid = g
        self.constraints.append(constraint)
        self._dirty = True

    # -------------------------------------------------------------- engine-facing protocol
    def _compute_data(self):
        if self._dirty:
            self._committed = self.store.compute_data()
            self._dirty = False
        return self._committed

    def _evaluate_committed(self):
        """totals and chi^2 of the committed histograms again (no histogram pass): what the reference's
        compute_data(update=False) returns when nothing moved; the scale factor may be refitted"""
        if self._dirty:
            return self._compute_data()
        if not self._resolved:
            raise RuntimeError("a move is staged; accept or reject it first")
        return self.store.finalize_data()

    def _before(self, relativeIndexes):
        # compute_before_move needs no device work of its own: the delta pass forms
        # after-minus-before in one sweep (PairDistributionConstraints.py:1053-1078 + :1095-1120)
        self._compute_data()
        # every constraint of a step is handed the SAME index object by the engine: seen once, nothing to compare
        # (the objects themselves are kept in _move_src, so an `is` can only match a live, unchanged argument of this step)
        if self._move is not None and self._move_src is not None and self._move_src[0] is relativeIndexes:
            return
        idx = np.ascontiguousarray(relativeIndexes, dtype=np.int32)
        if self._move is None or not np.array_equal(self._move[0], idx):
            self._move = (idx, None)
            self._move_src = (relativeIndexes, None)
            self._chi2 = None

    def _after(self, relativeIndexes, movedBoxCoordinates):
        if (self._chi2 is not None and self._move_src is not None and self._move_src[0] is relativeIndexes and
                self._move_src[1] is movedBoxCoordinates):
            return self._chi2                        # the second, third ... constraint of the step: already evaluated
        idx = np.ascontiguousarray(relativeIndexes, dtype=np.int32)
        moved = np.ascontiguousarray(movedBoxCoordinates, dtype=np.float32)
        same = (self._chi2 is not None and self._move is not None and self._move[1] is not None and
                np.array_equal(self._move[0], idx) and np.array_equal(self._move[1], moved))
        if not same:
            if not self._resolved:
                raise RuntimeError("previous move was neither accepted nor rejected")
            self._chi2 = self.store.propose(idx, moved)          # (a copy already)
            self._move = (idx, moved)
            self._move_src = (relativeIndexes, movedBoxCoordinates)
            self._resolved = False
        return self._chi2

    def _resolve(self, accept):
        if not self._resolved:
            (self.store.accept if accept else self.store.reject)()
            if accept:
                self._committed = self._chi2.copy()
                self.accepted += 1
            self._resolved = True
            self._move = None
            self._move_src = None
            self._chi2 = None

    # -------------------------------------------------------------- atom removal (Engine.py:3231-3276, :758-797)
    @property
    def numberOfAtoms(self):
        return self.store.numberOfAtoms

    def _amputate(self, relativeIndex):
        """compute_as_if_amputated of every registered constraint in one device evaluation"""
        rel = int(relativeIndex)
        if self._amputation is not None and self._amputation[0] == rel:
            return self._amputation[1]
        if not self._resolved or self._amputation is not None:
            raise RuntimeError("previous move was neither accepted nor rejected")
        self._compute_data()
        element = self.elements[int(self.elementsIndex[rel])]
        counts = dict(self.numberOfAtomsPerElement)
        counts[element] -= 1
        if counts[element] < 1:
            raise ValueError("Collecting last atom of any element type is not allowed")          # Engine.py:776
        rho0 = FLOAT_TYPE((self.numberOfAtoms - 1) / self.volume)                                 # PairDistributionConstraints.py:1198
        specs = [None] * self.store.n_models
        for c in self.constraints:
            c._amputationWeighting = c._weighting_for(counts)
            specs[c._model] = c._spec_with(counts, c._amputationWeighting, rho0)
        chi2 = self.store.propose_amputation(rel, specs, allow_fit=self.allowFittingScaleFactor).copy()
        self._amputation = (rel, chi2, counts)
        return chi2

    def _resolve_amputation(self, accept):
        if self._amputation is None:
            return
        rel, chi2, counts = self._amputation
        if accept:
            self.store.accept_amputation()
            self._committed = chi2.copy()
            self.accepted += 1
            self._pending_collect = (rel, counts)
        else:
            self.store.reject_amputation()
        self._amputation = None

    def _on_collector_collect_atom(self, relativeIndex=None):
        """Engine._on_collector_collect_atom (Engine.py:758-797) for the engine state this backend mirrors: the per-atom
        arrays lose the row, numberOfAtomsPerElement the atom, the number density follows in periodic systems only
        (:795-796); every model then gets the constants of that state (its weighting scheme is the one its
        accept_amputation adopted)."""
        rel, counts = self._pending_collect
        if relativeIndex is not None and int(relativeIndex) != rel:
            raise ValueError("collected atom %d is not the amputated one (%d)" % (int(relativeIndex), rel))
        self._pending_collect = None
        self.moleculesIndex = np.delete(self.moleculesIndex, rel, axis=0)
        self.elementsIndex = np.delete(self.elementsIndex, rel, axis=0)
        self.numberOfAtomsPerElement = counts
        if self.isPBC:
            self.numberDensity = FLOAT_TYPE(self.numberOfAtoms) / FLOAT_TYPE(self.volume)
        for c in self.constraints:
            self.store.set_model_constants(c._model, c._spec_with(counts, c.weighting, self.numberDensity))

    def close(self):
        self.store.close()


class _DeviceExperimentalConstraint(object):
    """Shared implementation of the five methods; subclasses fix the model kind."""
    KIND = None

    def __init__(self, backend, experimentalData, minDistance, maxDistance, bin, histSize, shellCenters, shellVolumes,
                 weighting, dataWeights=None, shapeArray=None, scaleFactor=1.0, qValues=None, adjustScaleFactor=(0, 0.8, 1.2),
                 shapeFuncParams=None, shapeWeighting=None, elementsWeight=None):
        self.backend = backend
        self.weighting = dict(weighting)                 # the constraint's weightingScheme ("A-B" -> float32)
        self.elementsWeight = None if elementsWeight is None else dict(elementsWeight)   # _elementsWeight (:486): only removals need it
        self.amputationStandardError = None
        self.experimentalData = np.ascontiguousarray(experimentalData, dtype=FLOAT_TYPE)
        self.minimumDistance = FLOAT_TYPE(minDistance)
        self.maximumDistance = FLOAT_TYPE(maxDistance)
        self.bin = FLOAT_TYPE(bin)
        self.histogramSize = int(histSize)
        self.shellCenters = np.ascontiguousarray(shellCenters, dtype=FLOAT_TYPE)
        self.shellVolumes = np.ascontiguousarray(shellVolumes, dtype=FLOAT_TYPE)
        self.standardError = None
        self.afterMoveStandardError = None
        self.tried = 0
        self.accepted = 0
        spec = ModelSpec(self.KIND, backend.elements, backend.numberOfAtomsPerElement, weighting, backend.volume,
                         backend.numberDensity, self.shellCenters, self.shellVolumes, self.experimentalData,
                         data_weights=dataWeights, shape_array=shapeArray, scale_factor=scaleFactor, q_values=qValues)
        self._spec = spec
        backend._register(self, (self.minimumDistance, self.maximumDistance, self.bin, self.histogramSize), spec)
        # shape function refreshed from the running configuration (set_shape_function_parameters with a dict,
        # PairDistributionConstraints.py:502-567): defaults and types as there
        self._shapeFuncParams, self._shapeUpdateFreq, self._lastShapeUpdate = None, 0, None
        self._shapeWeighting = shapeWeighting if shapeWeighting is not None else weighting
        self._shapeArray = None if shapeArray is None else np.ascontiguousarray(shapeArray, dtype=FLOAT_TYPE)
        if shapeFuncParams is not None:
            if self.KIND not in ("PDF", "PCF"):
                raise ValueError("only r-space constraints take shape function parameters")
            p = dict(shapeFuncParams)
            self._shapeFuncParams = {"rmin": FLOAT_TYPE(p.get("rmin", 0.00)), "rmax": p.get("rmax", None),
                                     "dr": FLOAT_TYPE(p.get("dr", 0.5)), "qmin": FLOAT_TYPE(p.get("qmin", 0.001)),
                                     "qmax": FLOAT_TYPE(p.get("qmax", 0.75)), "dq": FLOAT_TYPE(p.get("dq", 0.005))}
            self._shapeUpdateFreq = int(p.get("updateFreq", 1000))
        self.adjustScaleFactor = (int(adjustScaleFactor[0]), FLOAT_TYPE(adjustScaleFactor[1]), FLOAT_TYPE(adjustScaleFactor[2]))
        if self.adjustScaleFactor[0]:
            backend.store.set_adjust_scale_factor(self._model, *self.adjustScaleFactor)     # Core/Constraint.py:1196-1230

    # -- the reference's properties
    @property
    def data(self):
        """{"intra", "inter"} float32 (nEl,nEl,histSize) arrays, exported from the device on demand
        (save / plot / parity; never needed in the Monte-Carlo loop)."""
        self.backend._compute_data()
        intra, inter = self.backend.store.export_data(self._grid)
        return {"intra": intra, "inter": inter}

    @property
    def scaleFactor(self):
        """the constraint's scale factor (accept_move stores the fitted value, PairDistributionConstraints.py:1150)"""
        return self.backend.store.get_scale(self._model)[0]

    @property
    def fittedScaleFactor(self):
        """the scale factor the last evaluation used (the reference's _fittedScaleFactor)"""
        return self.backend.store.get_scale(self._model)[1]

    # -- optional tail of the total (PairDistributionConstraints.py:676-713; Core/Constraint.py:1160-1177)
    def set_window_function(self, windowFunction):
        """windowFunction: float32 array no longer than the data, or None; normalised here like the reference does"""
        if windowFunction is not None:
            windowFunction = np.array(windowFunction, dtype=FLOAT_TYPE)
            windowFunction /= np.sum(windowFunction)
        self.windowFunction = windowFunction
        self.backend.store.set_window_function(self._model, windowFunction)
        self.backend._dirty = True

    def set_multiframe_prior(self, multiframePrior, multiframeWeight):
        """total = multiframePrior + multiframeWeight * total; (None, None) switches it off"""
        self.backend.store.set_multiframe_prior(self._model, multiframePrior, 0.0 if multiframeWeight is None else multiframeWeight)
        self.backend._dirty = True

    # -- shape function (PairDistributionConstraints.py:316-374)
    def _update_shape_array(self):
        from . import shape
        b = self.backend
        p = self._shapeFuncParams
        coords = b.store.get_coords()
        rmax = p["rmax"]
        if rmax is None:
            rmax = shape.default_rmax(b.isPBC, b.basisVectors, coords)     # IBC: box coordinates are the real ones
        arr = shape.get_Gr_shape_function(self.shellCenters, coords, b.basisVectors, b.isPBC, b.moleculesIndex, b.elementsIndex,
                                          b.elements, b.numberOfAtomsPerElement, b.volume, self._shapeWeighting,
                                          qmin=p["qmin"], qmax=p["qmax"], dq=p["dq"], rmin=p["rmin"], rmax=rmax, dr=p["dr"])
        if self.KIND == "PCF":                                              # get_gr_shape_function (Collection.py:112-126)
            arr = arr / (FLOAT_TYPE(4.) * PI * b.numberDensity * self.shellCenters)
        self._shapeArray = arr
        b.store.set_shape(self._model, arr)

    def _reset_standard_error(self):
        chi2 = self.backend._evaluate_committed()
        self.standardError = FLOAT_TYPE(chi2[self._model])

    def runtime_initialize(self):
        """what Engine.run triggers before the first step (_runtime_initialize, :351-360)"""
        if self._shapeFuncParams is not None and self._shapeArray is None:
            self.backend._compute_data()
            self._update_shape_array()
        self.backend._compute_data()
        self._reset_standard_error()
        self._lastShapeUpdate = self.backend.accepted

    def runtime_on_step(self):
        """what Engine.run triggers before every step (_runtime_on_step, :362-374); returns True when the shape
        array was rebuilt (the caller then refreshes the engine's total standard error)"""
        if self._shapeUpdateFreq and self._shapeFuncParams is not None:
            acc = self.backend.accepted
            if self._lastShapeUpdate != acc and not (acc % self._shapeUpdateFreq):
                self._update_shape_array()
                self._reset_standard_error()
                self._lastShapeUpdate = acc
                return True
        return False

    def get_constraint_total(self, staged=False):
        """model total (G(r), g(r) or S(Q)) of the committed (or staged) state"""
        self.backend._compute_data()
        return self.backend.store.export_total(self._model, staged=staged)

    # -- atom removal (PairDistributionConstraints.py:1168-1238; PairCorrelationConstraints.py:394-462;
    #    StructureFactorConstraints.py:1098-1166)
    def _weighting_for(self, numberOfAtomsPerElement):
        """get_normalized_weighting(numbers, weights=self._elementsWeight) cast to FLOAT_TYPE (:1190-1192)"""
        if self.elementsWeight is None:
            raise ValueError("removing atoms needs the constraint's elementsWeight (element -> weight)")
        return faber_ziman_weights(numberOfAtomsPerElement, self.elementsWeight)

    def _spec_with(self, numberOfAtomsPerElement, weighting, rho0):
        """this constraint's model constants for another composition / number density"""
        o = self._spec
        return ModelSpec(o.kind, o.elements, numberOfAtomsPerElement, weighting, o.volume, rho0, o.shell_centers, o.shell_volumes,
                         o.experimental, data_weights=o.data_weights, shape_array=o.shape_array, scale_factor=o.scale_factor,
                         gr2sq=o.gr2sq, sq_exact=o.sq_exact)

    def compute_as_if_amputated(self, realIndex, relativeIndex):
        chi2 = self.backend._amputate(np.asarray(relativeIndex).ravel()[0])
        self.amputationStandardError = FLOAT_TYPE(chi2[self._model])

    def accept_amputation(self, realIndex, relativeIndex):
        self.backend._resolve_amputation(True)
        self.weighting = self._amputationWeighting
        self.standardError = self.amputationStandardError
        self.amputationStandardError = None

    def reject_amputation(self, realIndex, relativeIndex):
        self.backend._resolve_amputation(False)
        self.amputationStandardError = None

    def _on_collector_collect_atom(self, realIndex):
        pass                                             # like the reference's: the engine-level hook does the work

    def set_data(self, data):
        """resume from saved data["intra"] / data["inter"] (Constraint.set_data; what the repository holds of this
        constraint, Core/Constraint.py:275-288) instead of a compute_data pass"""
        b = self.backend
        b.store.import_data(self._grid, data["intra"], data["inter"])
        b._imported = getattr(b, "_imported", set()) | {self._grid}
        if b._imported >= set(b._grids.values()):          # every grid of the store has its counts: totals and chi^2
            b._committed = b.store.finalize_data()
            b._dirty = False
            b._imported = set()
            for c in b.constraints:
                c.standardError = FLOAT_TYPE(b._committed[c._model])

    # -- the five methods (Core/Constraint.py:732-748)
    def compute_data(self, update=True):
        if update:
            self.backend._dirty = True
            chi2 = self.backend._compute_data()
        else:
            chi2 = self.backend._evaluate_committed()
        if update:
            self.standardError = FLOAT_TYPE(chi2[self._model])
        return self.data, FLOAT_TYPE(chi2[self._model])

    def compute_before_move(self, realIndexes, relativeIndexes):
        self.backend._before(relativeIndexes)
        if self.standardError is None:
            self.standardError = FLOAT_TYPE(self.backend._committed[self._model])

    def compute_after_move(self, realIndexes, relativeIndexes, movedBoxCoordinates):
        chi2 = self.backend._after(relativeIndexes, movedBoxCoordinates)
        self.afterMoveStandardError = FLOAT_TYPE(chi2[self._model])
        self.tried += 1

    def accept_move(self, realIndexes, relativeIndexes):
        self.backend._resolve(True)
        self.standardError = self.afterMoveStandardError
        self.afterMoveStandardError = None
        self.accepted += 1

    def reject_move(self, realIndexes, relativeIndexes):
        self.backend._resolve(False)
        self.afterMoveStandardError = None


class DevicePairDistributionConstraint(_DeviceExperimentalConstraint):
    """G(r) constraint; experimentalData is the (n,2) [r, G(r)] array of the reference
    (set_experimental_data / set_limits, PairDistributionConstraints.py:700-768)."""
    KIND = "PDF"

    def __init__(self, backend, experimentalData, weighting, dataWeights=None, shapeArray=None, scaleFactor=1.0,
                 adjustScaleFactor=(0, 0.8, 1.2), shapeFuncParams=None, shapeWeighting=None, elementsWeight=None):
        exp = np.ascontiguousarray(experimentalData, dtype=FLOAT_TYPE)
        r = exp[:, 0]
        b = FLOAT_TYPE(r[1] - r[0])                                            # :726
        rmin = FLOAT_TYPE(r[0] - b / 2.)                                       # :747
        rmax = FLOAT_TYPE(r[-1] + b / 2.)                                      # :748
        edges = np.array([x - b / 2. for x in r] + [r[-1] + b / 2.], dtype=FLOAT_TYPE)   # :751-753
        hs = len(edges) - 1
        super(DevicePairDistributionConstraint, self).__init__(
            backend, exp[:, 1], rmin, rmax, b, hs, np.array(r, dtype=FLOAT_TYPE), shell_volumes_from_edges(edges),
            weighting, dataWeights, shapeArray, scaleFactor, adjustScaleFactor=adjustScaleFactor,
            shapeFuncParams=shapeFuncParams, shapeWeighting=shapeWeighting, elementsWeight=elementsWeight)


class DevicePairCorrelationConstraint(DevicePairDistributionConstraint):
    """g(r) constraint (PairCorrelationConstraints.py:126-169)."""
    KIND = "PCF"


class DeviceStructureFactorConstraint(_DeviceExperimentalConstraint):
    """S(Q) constraint: experimentalData is (m,2) [Q, S(Q)].  The r-grid follows the reference's defaults
    (StructureFactorConstraints.py): rmin=None -> 2 pi / Qmax (:511-512); dr=None -> 2 pi / Qmax rounded DOWN to one
    decimal (:560-565); rmax given -> edges = arange(rmin, rmax + dr, dr) (:341), rmax=None -> edges = arange(rmin,
    half the shortest basis vector, dr) (:337-339; the reference uses 2 pi / dQ only while no engine is attached)."""
    KIND = "SQ"

    def __init__(self, backend, experimentalData, weighting, rmin=None, rmax=None, dr=None, dataWeights=None, scaleFactor=1.0,
                 adjustScaleFactor=(0, 0.8, 1.2), elementsWeight=None):
        exp = np.ascontiguousarray(experimentalData, dtype=FLOAT_TYPE)
        qmax = exp[-1, 0]
        minimumDistance = FLOAT_TYPE(2. * PI / qmax) if rmin is None else FLOAT_TYPE(rmin)
        if dr is None:
            b = 2. * PI / qmax
            rb = round(b, 1)
            if rb > b:
                rb -= 0.1
            b = FLOAT_TYPE(rb)
        else:
            b = FLOAT_TYPE(dr)
        if rmax is None:
            half = np.min([np.linalg.norm(v) / 2. for v in backend.basisVectors])
            edges = np.arange(minimumDistance, half, b).astype(FLOAT_TYPE)
        else:
            edges = np.arange(minimumDistance, FLOAT_TYPE(rmax) + b, b).astype(FLOAT_TYPE)
        centers = (edges[0:-1] + edges[1:]) / FLOAT_TYPE(2.)                   # :346
        hs = len(edges) - 1
        super(DeviceStructureFactorConstraint, self).__init__(
            backend, exp[:, 1], edges[0], edges[-1], b, hs, centers, shell_volumes_from_edges(edges),
            weighting, dataWeights, None, scaleFactor, qValues=exp[:, 0], adjustScaleFactor=adjustScaleFactor,
            elementsWeight=elementsWeight)


class DeviceReducedStructureFactorConstraint(DeviceStructureFactorConstraint):
    """S(Q)-1 normalisation (StructureFactorConstraints.py:1235-1260)."""
    KIND = "RSQ"


_KIND_CLASS = {}


def make_device_constraint(backend, kind, experimental, minDistance, maxDistance, bin, histSize, shellCenters, shellVolumes,
                           weighting, dataWeights=None, shapeArray=None, scaleFactor=1.0, qValues=None,
                           adjustScaleFactor=(0, 0.8, 1.2), shapeFuncParams=None, shapeWeighting=None, elementsWeight=None):
    """Build a device constraint from the quantities a reference constraint has already derived
    (limits, bin, histogram size, shell arrays, weighting scheme) -- what the subclass recipe of
    INTEGRATION.md hands over."""
    if not _KIND_CLASS:
        for name in ("PDF", "PCF", "SQ", "RSQ"):
            _KIND_CLASS[name] = type("Device%sConstraint" % name, (_DeviceExperimentalConstraint,), {"KIND": name})
    return _KIND_CLASS[kind](backend, experimental, minDistance, maxDistance, bin, histSize, shellCenters, shellVolumes,
                             weighting, dataWeights, shapeArray, scaleFactor, qValues, adjustScaleFactor, shapeFuncParams,
                             shapeWeighting, elementsWeight)
