"""DeviceStore -- Python face of the stateful C ABI (``frmc_store_*``, ``frmc_propose`` ...).

One store holds the system (fractional coordinates, molecule and element indexes) on one
GPU in the element-sorted 16-byte-record layout (csrc/layout.h), any number of r-grids
(each with its running integer histograms) and the models (constraints) evaluated on them.
"""
import ctypes

import numpy as np

from . import _lib as L
from .model import ModelSpec

_F32, _I32 = np.float32, np.int32


class DeviceStore(object):
    """Device-resident coordinate store + running pair histograms.

    :Parameters:
        #. boxCoords (float32 (N,3)): engine.boxCoordinates (fractional; Cartesian when not isPBC).
        #. basis (float32 (3,3)): engine.basisVectors.
        #. isPBC (bool): engine.isPBC.
        #. moleculeIndex, elementIndex (int32 (N,)): engine.moleculesIndex / elementsIndex.
        #. numberOfElements (int): engine.numberOfElements.
        #. device (None, int): CUDA device; default from $FULLRMC_B200_DEVICE / $LOCAL_RANK / 0.
    """

    def __init__(self, boxCoords, basis, isPBC, moleculeIndex, elementIndex, numberOfElements, device=None):
        self._lib = L.load_library()
        self._handle = None
        coords = L.as_array(boxCoords, "boxCoords", _F32, 2)
        basis = L.as_array(basis, "basis", _F32, 2)
        mol = L.as_array(moleculeIndex, "moleculeIndex", _I32, 1)
        el = L.as_array(elementIndex, "elementIndex", _I32, 1)
        if coords.shape[1] != 3 or basis.shape != (3, 3):
            raise ValueError("boxCoords must be (N,3) and basis (3,3)")
        if mol.shape[0] != coords.shape[0] or el.shape[0] != coords.shape[0]:
            raise ValueError("moleculeIndex/elementIndex length must equal the number of atoms")
        self.device = L.device_index() if device is None else int(device)
        self.numberOfAtoms = coords.shape[0]
        self.numberOfElements = int(numberOfElements)
        self.isPBC = bool(isPBC)
        h = self._lib.frmc_store_create(self.device, coords.shape[0], L.ptr(coords, L.c_f32p), L.ptr(basis, L.c_f32p),
                                        int(self.isPBC), L.ptr(mol, L.c_i32p), L.ptr(el, L.c_i32p),
                                        self.numberOfElements)
        if not h:
            raise L.FullrmcB200Error("frmc_store_create: " + L.last_error())
        self._handle = ctypes.c_void_p(h)
        self._grids = []          # (rmin, rmax, bin, hs)
        self._models = []         # ModelSpec
        self._keep = []           # arrays referenced by descriptors during frmc_model_add
        self._chi2 = np.zeros(8, dtype=_F32)
        self._chi2_ptr = L.ptr(self._chi2, L.c_f32p)
        self._step_fn = self._lib.frmc_step
        # frmc_propose once more with plain addresses for the array arguments (no ctypes cast per call: see step())
        self._propose_fn = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            L.c_f32p)(("frmc_propose", self._lib))

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if self._handle is not None:
            self._lib.frmc_store_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def stream(self):
        """cudaStream_t (int) the store launches on."""
        return int(self._lib.frmc_store_stream(self._handle) or 0)

    # ------------------------------------------------------------------ configuration
    def add_grid(self, minDistance, maxDistance, bin, histSize):
        g = L.check(self._lib.frmc_grid_add(self._handle, float(_F32(minDistance)), float(_F32(maxDistance)),
                                            float(_F32(bin)), int(histSize)), "add_grid")
        self._grids.append((float(minDistance), float(maxDistance), float(bin), int(histSize)))
        return g

    def add_model(self, grid, spec):
        if not isinstance(spec, ModelSpec):
            raise TypeError("spec must be a ModelSpec")
        desc, keep = spec.build_desc()
        m = L.check(self._lib.frmc_model_add(self._handle, int(grid), ctypes.byref(desc)), "add_model")
        self._models.append(spec)
        return m

    def set_scale(self, model, scale):
        L.check(self._lib.frmc_model_set_scale(self._handle, int(model), float(_F32(scale))), "set_scale")

    def set_shape(self, model, shape_array):
        """replace (None: drop) the shape-function array of an r-space model; follow with finalize_data()"""
        if shape_array is None:
            L.check(self._lib.frmc_model_set_shape(self._handle, int(model), None), "set_shape")
            return
        a = np.ascontiguousarray(shape_array, dtype=_F32)
        L.check(self._lib.frmc_model_set_shape(self._handle, int(model), L.ptr(a, L.c_f32p)), "set_shape")

    def set_window_function(self, model, window):
        """normalised window function convolved with the model total (set_window_function,
        PairDistributionConstraints.py:676-713); None switches it off"""
        if window is None:
            L.check(self._lib.frmc_model_set_window(self._handle, int(model), None, 0), "set_window_function")
            return
        w = np.ascontiguousarray(window, dtype=_F32)
        L.check(self._lib.frmc_model_set_window(self._handle, int(model), L.ptr(w, L.c_f32p), int(w.shape[0])), "set_window_function")

    def set_multiframe_prior(self, model, prior, weight):
        """total = prior + weight * total (Core/Constraint.py:1160-1177); prior None switches it off"""
        if prior is None:
            L.check(self._lib.frmc_model_set_multiframe_prior(self._handle, int(model), None, 0.0), "set_multiframe_prior")
            return
        a = np.ascontiguousarray(prior, dtype=_F32)
        L.check(self._lib.frmc_model_set_multiframe_prior(self._handle, int(model), L.ptr(a, L.c_f32p), float(_F32(weight))),
                "set_multiframe_prior")

    def set_adjust_scale_factor(self, model, frequency, minimum, maximum):
        """ExperimentalConstraint.set_adjust_scale_factor (Core/Constraint.py:1160-1177): refit the scale factor in
        every evaluation made while accepted % frequency == 0, clipped to [minimum, maximum]; 0 switches it off."""
        L.check(self._lib.frmc_model_set_adjust(self._handle, int(model), int(frequency), float(_F32(minimum)),
                                                float(_F32(maximum))), "set_adjust_scale_factor")

    def get_scale(self, model):
        """(the model's scale factor, the scale factor the last evaluation used)"""
        a = ctypes.c_float(0.0); b = ctypes.c_float(0.0)
        L.check(self._lib.frmc_model_get_scale(self._handle, int(model), ctypes.byref(a), ctypes.byref(b)), "get_scale")
        return _F32(a.value), _F32(b.value)

    def set_persistent(self, on=True):
        """Keep one cooperative kernel resident across a run of propose/step calls (include/fullrmc_b200.h:
        frmc_store_set_persistent); results are identical, the per-move kernel launch disappears."""
        L.check(self._lib.frmc_store_set_persistent(self._handle, int(bool(on))), "set_persistent")

    def persistent_stats(self):
        """(persistent kernels started, proposals they served)"""
        a = ctypes.c_uint64(0); b = ctypes.c_uint64(0)
        L.check(self._lib.frmc_store_persistent_stats(self._handle, ctypes.byref(a), ctypes.byref(b)), "persistent_stats")
        return int(a.value), int(b.value)

    def set_accepted(self, accepted):
        """re-base the store's count of accepted moves on engine.accepted (it drives the refit schedule)"""
        L.check(self._lib.frmc_store_set_accepted(self._handle, int(accepted)), "set_accepted")

    @property
    def n_models(self):
        return len(self._models)

    # ------------------------------------------------------------------ compute_data
    def compute_data(self):
        """Full histogram of every grid + totals; returns float32 chi^2 per model."""
        L.check(self._lib.frmc_compute_data(self._handle, L.ptr(self._chi2, L.c_f32p)), "compute_data")
        return self._chi2[:self.n_models].copy()

    def compute_data_shard(self, shard, nshards):
        """This rank's share of the tile work list only (device-resident partial counts)."""
        L.check(self._lib.frmc_compute_data_shard(self._handle, int(shard), int(nshards)), "compute_data_shard")

    def counts_pointer(self, grid):
        """(device pointer, number of int64 cells) of the grid's committed counts [2][nEl*nEl][hs]."""
        n = ctypes.c_int64(0)
        p = self._lib.frmc_grid_counts_ptr(self._handle, int(grid), ctypes.byref(n))
        if not p:
            raise L.FullrmcB200Error("grid_counts_ptr: " + L.last_error())
        return int(p), int(n.value)

    def finalize_data(self):
        L.check(self._lib.frmc_finalize_data(self._handle, L.ptr(self._chi2, L.c_f32p)), "finalize_data")
        return self._chi2[:self.n_models].copy()

    # ------------------------------------------------------------------ per-move path
    def propose(self, indexes, movedBoxCoordinates):
        """compute_before_move + compute_after_move in one pass; returns chi^2-after per model."""
        idx, moved = indexes, movedBoxCoordinates
        if not (type(idx) is np.ndarray and idx.dtype == _I32 and idx.flags.c_contiguous):
            idx = np.ascontiguousarray(idx, dtype=_I32)
        if not (type(moved) is np.ndarray and moved.dtype == _F32 and moved.flags.c_contiguous):
            moved = np.ascontiguousarray(moved, dtype=_F32)
        if moved.shape != (idx.shape[0], 3):
            raise ValueError("movedBoxCoordinates must be (k,3)")
        rc = self._propose_fn(self._handle, idx.__array_interface__["data"][0], idx.shape[0], moved.__array_interface__["data"][0],
                              self._chi2_ptr)
        if rc < 0:
            L.check(rc, "propose")
        return self._chi2[:len(self._models)].copy()

    def step(self, previous, indexes, movedBoxCoordinates):
        """Resolve the staged proposal (previous: True accept / False reject / None nothing staged)
        and evaluate the next one with a single call into the library.  Returns a VIEW of the
        chi^2 buffer (valid until the next call); this is the lean entry point for tight loops."""
        idx, moved = indexes, movedBoxCoordinates
        if not (type(idx) is np.ndarray and idx.dtype == _I32 and idx.flags.c_contiguous):
            idx = np.ascontiguousarray(idx, dtype=_I32)
        if not (type(moved) is np.ndarray and moved.dtype == _F32 and moved.flags.c_contiguous):
            moved = np.ascontiguousarray(moved, dtype=_F32)
        k = idx.shape[0]
        if moved.size != 3 * k:
            raise ValueError("movedBoxCoordinates must be (k,3)")
        prev = -1 if previous is None else (1 if previous else 0)
        rc = self._step_fn(self._handle, prev, idx.__array_interface__["data"][0], k, moved.__array_interface__["data"][0],
                           self._chi2_ptr)
        if rc < 0:
            L.check(rc, "step")
        return self._chi2[:len(self._models)]

    def run_batch(self, indexes, movedBoxCoordinates, total, rand, tolerance=0.0, group_sizes=None, variance_squared=None):
        """A run of proposals tried in order with the engine's acceptance rule, resolved on the device
        (include/fullrmc_b200.h: frmc_run_batch; Engine.py:3302-3338).

        :Parameters:
            #. indexes (int32 (A,)): the moved atoms of all proposals back to back.
            #. movedBoxCoordinates (float32 (A,3)): their proposed box coordinates.
            #. total (float): engine.totalStandardError before the run.
            #. rand (float32 (n,)): pre-drawn uniform numbers, one consumed per worse proposal.
            #. tolerance (float): engine.tolerance.
            #. group_sizes (None, int32 (n,)): atoms per proposal (None: one each).
            #. variance_squared (None, float32 (n_models,)): constraint.varianceSquared (None: 1).

        :Returns: dict with chi2 (n, n_models), decisions (n,) int32 (0 rejected / 1 accepted / 2 tolerated),
            total (float32), rand_used (int), device_ms (float).
        """
        idx = np.ascontiguousarray(indexes, dtype=_I32).ravel()
        moved = np.ascontiguousarray(movedBoxCoordinates, dtype=_F32)
        if moved.shape != (idx.shape[0], 3):
            raise ValueError("movedBoxCoordinates must be (A,3)")
        if group_sizes is None:
            n, gs = idx.shape[0], None
        else:
            gs = np.ascontiguousarray(group_sizes, dtype=_I32).ravel()
            n = gs.shape[0]
            if int(gs.sum()) != idx.shape[0]:
                raise ValueError("group_sizes must add up to the number of indexes")
        rnd = np.ascontiguousarray(rand, dtype=_F32).ravel()
        if rnd.shape[0] < n:
            raise ValueError("rand must hold one number per proposal")
        var = None if variance_squared is None else np.ascontiguousarray(variance_squared, dtype=_F32).ravel()
        if var is not None and var.shape[0] != self.n_models:
            raise ValueError("variance_squared must hold one value per model")
        chi2 = np.zeros((n, self.n_models), dtype=_F32)
        dec = np.zeros(n, dtype=_I32)
        tot = ctypes.c_float(float(_F32(total)))
        used = ctypes.c_int32(0)
        ms = ctypes.c_double(0.0)
        L.check(self._lib.frmc_run_batch(self._handle, n, L.ptr(gs, L.c_i32p), L.ptr(idx, L.c_i32p), L.ptr(moved, L.c_f32p),
                                         L.ptr(var, L.c_f32p), float(_F32(tolerance)), L.ptr(rnd, L.c_f32p), ctypes.byref(tot),
                                         L.ptr(chi2, L.c_f32p), L.ptr(dec, L.c_i32p), ctypes.byref(used), ctypes.byref(ms)),
                "run_batch")
        return {"chi2": chi2, "decisions": dec, "total": _F32(tot.value), "rand_used": int(used.value), "device_ms": float(ms.value)}

    def batch_stats(self):
        """(batch launches, evaluation rounds inside them, proposals resolved)"""
        a = ctypes.c_uint64(0); b = ctypes.c_uint64(0); c = ctypes.c_uint64(0)
        L.check(self._lib.frmc_store_batch_stats(self._handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "batch_stats")
        return int(a.value), int(b.value), int(c.value)

    def committed_chi2(self):
        out = np.zeros(8, dtype=_F32)
        L.check(self._lib.frmc_store_committed_chi2(self._handle, L.ptr(out, L.c_f32p)), "committed_chi2")
        return out[:self.n_models].copy()

    def replay_proposal(self, reps):
        """Average device time (ms) of the staged proposal's pipeline over `reps` back-to-back launches."""
        ms = ctypes.c_double(0.0)
        L.check(self._lib.frmc_store_replay_proposal(self._handle, int(reps), ctypes.byref(ms)), "replay_proposal")
        return float(ms.value)

    def accept(self):
        L.check(self._lib.frmc_accept(self._handle), "accept")

    def reject(self):
        L.check(self._lib.frmc_reject(self._handle), "reject")

    # ------------------------------------------------------------------ distance-constraint pre-filter on the store
    def distance_add(self, typesIndex, numberOfTypes, lowerLimit, upperLimit, interMolecular=True, intraMolecular=True,
                     countWithinLimits=True, reduceDistanceToUpper=False, reduceDistanceToLower=False, reduceDistance=False):
        """register a molecular distance constraint on this store (include/fullrmc_b200.h: frmc_store_distance_add);
        the limit arrays are the reference's [nT,nT,1] arrays indexed [type_i, type_a]; returns its id"""
        from .Core.atomic_distances import _flags
        t = np.ascontiguousarray(typesIndex, dtype=_I32)
        nT = int(numberOfTypes)
        lo = np.ascontiguousarray(lowerLimit, dtype=_F32).reshape(-1)
        up = np.ascontiguousarray(upperLimit, dtype=_F32).reshape(-1)
        if t.shape[0] != self.numberOfAtoms or lo.shape[0] != nT * nT or up.shape[0] != nT * nT:
            raise ValueError("typesIndex must have one entry per atom and the limits numberOfTypes^2 entries")
        flags = _flags(interMolecular, intraMolecular, countWithinLimits, reduceDistanceToUpper, reduceDistanceToLower, reduceDistance)
        cid = L.check(self._lib.frmc_store_distance_add(self._handle, L.ptr(t, L.c_i32p), nT, L.ptr(lo, L.c_f32p), L.ptr(up, L.c_f32p),
                                                        flags), "distance_add")
        self._dist = getattr(self, "_dist", {})
        counts, sums = np.zeros((4, 2, nT, nT, 1), dtype=_I32), np.zeros((4, 2, nT, nT, 1), dtype=_F32)
        self._dist[cid] = (counts, sums, counts.__array_interface__["data"][0], sums.__array_interface__["data"][0])
        return cid

    def distance_move(self, cid, indexes, movedBoxCoordinates):
        """One pass over the resident atoms for a move of a registered distance constraint: returns (counts, sums), each
        (4, 2, nT, nT, 1) = (M before, F before, M after, F after) x (intra, inter): M = multiple_atomic_distances_coords of
        the group against all atoms, F = full_atomic_distances_coords of the group alone.  Views of reused buffers."""
        idx, moved = indexes, movedBoxCoordinates
        if not (type(idx) is np.ndarray and idx.dtype == _I32 and idx.flags.c_contiguous):
            idx = np.ascontiguousarray(idx, dtype=_I32)
        if not (type(moved) is np.ndarray and moved.dtype == _F32 and moved.flags.c_contiguous):
            moved = np.ascontiguousarray(moved, dtype=_F32)
        if moved.size != 3 * idx.shape[0]:
            raise ValueError("movedBoxCoordinates must be (k,3)")
        counts, sums, pc, ps = self._dist[cid]
        rc = self._lib.frmc_store_distance_move(self._handle, cid, idx.__array_interface__["data"][0], idx.shape[0],
                                                moved.__array_interface__["data"][0], pc, ps)
        if rc < 0:
            L.check(rc, "distance_move")
        return counts, sums

    # ------------------------------------------------------------------ coordination-number pre-filter on the store
    def coordination_add(self, coresIndexes, shellsIndexes, lowerShells, upperShells):
        """register the definitions of a coordination-number constraint on this store (include/fullrmc_b200.h:
        frmc_store_coordination_add): per definition the core and shell atom lists and the shell bounds; returns its id"""
        ndef = len(coresIndexes)
        if not (len(shellsIndexes) == len(lowerShells) == len(upperShells) == ndef):
            raise ValueError("every per-definition argument needs one entry per shell definition (%d)" % ndef)

        def flat(lists):
            arrays = [np.ascontiguousarray(a, dtype=_I32).reshape(-1) for a in lists]
            off = np.zeros(ndef + 1, np.int64)
            off[1:] = np.cumsum([a.shape[0] for a in arrays])
            idx = np.concatenate(arrays) if off[-1] else np.zeros(1, _I32)
            return off, np.ascontiguousarray(idx, dtype=_I32)

        coff, cidx = flat(coresIndexes)
        soff, sidx = flat(shellsIndexes)
        lo = np.array([_F32(x) for x in lowerShells], dtype=_F32)
        up = np.array([_F32(x) for x in upperShells], dtype=_F32)
        cid = L.check(self._lib.frmc_store_coordination_add(self._handle, ndef, L.ptr(coff, L.c_i64p), L.ptr(cidx, L.c_i32p),
                                                            L.ptr(soff, L.c_i64p), L.ptr(sidx, L.c_i32p), L.ptr(lo, L.c_f32p),
                                                            L.ptr(up, L.c_f32p)), "coordination_add")
        self._coord = getattr(self, "_coord", {})
        counts = np.zeros((2, ndef), dtype=_I32)
        self._coord[cid] = (counts, counts.__array_interface__["data"][0])
        return cid

    def coordination_move(self, cid, indexes, movedBoxCoordinates):
        """One launch over the resident atoms for a move of a registered coordination-number constraint: int32 (2, nDef) =
        what multi_atoms_coord_number_coords adds to coordNumData for the group before and after the move.  A view of a
        reused buffer."""
        idx, moved = indexes, movedBoxCoordinates
        if not (type(idx) is np.ndarray and idx.dtype == _I32 and idx.flags.c_contiguous):
            idx = np.ascontiguousarray(idx, dtype=_I32)
        if not (type(moved) is np.ndarray and moved.dtype == _F32 and moved.flags.c_contiguous):
            moved = np.ascontiguousarray(moved, dtype=_F32)
        if moved.size != 3 * idx.shape[0]:
            raise ValueError("movedBoxCoordinates must be (k,3)")
        if cid not in getattr(self, "_coord", {}):
            raise ValueError("unknown coordination constraint %r (coordination_add returns the id)" % (cid,))
        counts, pc = self._coord[cid]
        rc = self._lib.frmc_store_coordination_move(self._handle, cid, idx.__array_interface__["data"][0], idx.shape[0],
                                                    moved.__array_interface__["data"][0], pc)
        if rc < 0:
            L.check(rc, "coordination_move")
        return counts

    def move_atoms(self, indexes, movedBoxCoordinates):
        """apply an accepted move on a store without histogram models (with models, accept() does it)"""
        idx = np.ascontiguousarray(indexes, dtype=_I32)
        moved = np.ascontiguousarray(movedBoxCoordinates, dtype=_F32)
        L.check(self._lib.frmc_store_move_atoms(self._handle, L.ptr(idx, L.c_i32p), idx.shape[0], L.ptr(moved, L.c_f32p)), "move_atoms")

    # ------------------------------------------------------------------ device-generated runs of moves
    def set_real_coords(self, realCoordinates=None, reciprocalBasisVectors=None):
        """engine.realCoordinates and engine.reciprocalBasisVectors (both None for a non-periodic store)"""
        if not self.isPBC:
            L.check(self._lib.frmc_store_set_real_coords(self._handle, None, None), "set_real_coords")
            return
        real = L.as_array(realCoordinates, "realCoordinates", _F32, 2)
        rb = L.as_array(reciprocalBasisVectors, "reciprocalBasisVectors", _F32, 2)
        if real.shape != (self.numberOfAtoms, 3) or rb.shape != (3, 3):
            raise ValueError("realCoordinates must be (N,3) and reciprocalBasisVectors (3,3)")
        L.check(self._lib.frmc_store_set_real_coords(self._handle, L.ptr(real, L.c_f32p), L.ptr(rb, L.c_f32p)), "set_real_coords")

    def get_real_coords(self):
        out = np.empty((self.numberOfAtoms, 3), dtype=_F32)
        L.check(self._lib.frmc_store_get_real_coords(self._handle, L.ptr(out, L.c_f32p)), "get_real_coords")
        return out

    def set_groups(self, groups=None, offsets=None, indexes=None):
        """the engine's groups: a list of atom-index lists (engine.groups[g].indexes), None for one group per atom
        (Engine.set_groups(None)), or the flat form offsets (G+1,) / indexes"""
        if offsets is not None:
            off = np.ascontiguousarray(offsets, dtype=_I32)
            idx = np.ascontiguousarray(indexes, dtype=_I32)
        elif groups is None:
            off = np.arange(self.numberOfAtoms + 1, dtype=_I32)
            idx = np.arange(self.numberOfAtoms, dtype=_I32)
        else:
            off = np.zeros(len(groups) + 1, dtype=_I32)
            off[1:] = np.cumsum([len(g) for g in groups])
            idx = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=_I32).ravel() for g in groups]), dtype=_I32)
        L.check(self._lib.frmc_store_set_groups(self._handle, off.shape[0] - 1, L.ptr(off, L.c_i32p), L.ptr(idx, L.c_i32p)), "set_groups")

    def run_generated(self, n, seed, first_counter, amplitude, total, tolerance=0.0, variance_squared=None):
        """n Metropolis steps generated, evaluated, decided and applied on the device (include/fullrmc_b200.h:
        frmc_run_generated; the random-number contract is fullrmc_b200/rng.py).  amplitude: number or (min, max) like
        TranslationGenerator.  Returns dict with chi2 (n, n_models), decisions (n,), groups (n,), rand (n,) the
        acceptance numbers, total (float32), device_ms."""
        amp = (0.0, float(amplitude)) if np.isscalar(amplitude) else (float(amplitude[0]), float(amplitude[1]))
        var = None if variance_squared is None else np.ascontiguousarray(variance_squared, dtype=_F32).ravel()
        chi2 = np.zeros((n, self.n_models), dtype=_F32)
        dec = np.zeros(n, dtype=_I32)
        grp = np.zeros(n, dtype=_I32)
        rnd = np.zeros(n, dtype=_F32)
        tot = ctypes.c_float(float(_F32(total)))
        ms = ctypes.c_double(0.0)
        L.check(self._lib.frmc_run_generated(self._handle, int(n), int(seed), int(first_counter), float(_F32(amp[0])), float(_F32(amp[1])),
                                             L.ptr(var, L.c_f32p), float(_F32(tolerance)), ctypes.byref(tot), L.ptr(chi2, L.c_f32p),
                                             L.ptr(dec, L.c_i32p), L.ptr(grp, L.c_i32p), L.ptr(rnd, L.c_f32p), ctypes.byref(ms)),
                "run_generated")
        return {"chi2": chi2, "decisions": dec, "groups": grp, "rand": rnd, "total": _F32(tot.value), "device_ms": float(ms.value)}

    # ------------------------------------------------------------------ atom removal, persisted state
    def _constants(self, spec):
        pa, pb, pw, pD = spec.pair_table()
        return pw, pD, spec.prefactor()

    def propose_amputation(self, index, specs=None, allow_fit=False):
        """compute_as_if_amputated for every model (include/fullrmc_b200.h: frmc_propose_amputation): chi^2 per model of
        the system without atom `index`.  specs: one ModelSpec per model holding the constants of that system (weighting
        scheme, numberOfAtomsPerElement and number density with the atom gone), or None to keep the models' own."""
        descs, keep = None, []
        if specs is not None:
            if len(specs) != self.n_models:
                raise ValueError("one ModelSpec per model")
            descs = (L.AmputationDesc * self.n_models)()
            for i, spec in enumerate(specs):
                if spec is None:
                    continue
                pw, pD, pref = self._constants(spec)
                keep += [pw, pD, pref]
                descs[i].pair_w = L.ptr(pw, L.c_f32p); descs[i].pair_D = L.ptr(pD, L.c_f32p); descs[i].prefactor = L.ptr(pref, L.c_f32p)
        L.check(self._lib.frmc_propose_amputation(self._handle, int(index), descs, int(bool(allow_fit)), L.ptr(self._chi2, L.c_f32p)),
                "propose_amputation")
        return self._chi2[:self.n_models].copy()

    def accept_amputation(self, specs=None):
        """accept_amputation + Engine._on_collector_collect_atom: the atom leaves the store; specs (one ModelSpec per model
        or None) are the constants the models use from now on."""
        L.check(self._lib.frmc_accept_amputation(self._handle), "accept_amputation")
        self.numberOfAtoms = int(self._lib.frmc_store_n_atoms(self._handle))
        if specs is not None:
            for i, spec in enumerate(specs):
                if spec is not None:
                    self.set_model_constants(i, spec)

    def reject_amputation(self):
        L.check(self._lib.frmc_reject_amputation(self._handle), "reject_amputation")

    def set_model_constants(self, model, spec):
        """replace a model's weighting scheme / D_ij / 4 pi r rho0 by those of `spec`"""
        pw, pD, pref = self._constants(spec)
        L.check(self._lib.frmc_model_set_constants(self._handle, int(model), L.ptr(pw, L.c_f32p), L.ptr(pD, L.c_f32p),
                                                   L.ptr(pref, L.c_f32p)), "set_model_constants")
        self._models[model] = spec

    def import_data(self, grid, hintra, hinter):
        """Resume from saved data["intra"] / data["inter"] (float32 (nEl,nEl,hs)) instead of a full-histogram pass;
        follow with finalize_data()."""
        a = np.ascontiguousarray(hintra, dtype=_F32)
        b = np.ascontiguousarray(hinter, dtype=_F32)
        hs = self._grids[grid][3]
        if a.shape != (self.numberOfElements, self.numberOfElements, hs) or b.shape != a.shape:
            raise ValueError("saved histograms must be (nEl, nEl, histSize)")
        L.check(self._lib.frmc_import_data(self._handle, int(grid), L.ptr(a, L.c_f32p), L.ptr(b, L.c_f32p)), "import_data")

    # ------------------------------------------------------------------ export
    def export_data(self, grid=0):
        """The reference's data["intra"], data["inter"]: float32 (nEl,nEl,hs) arrays."""
        hs = self._grids[grid][3]
        nEl = self.numberOfElements
        hintra = np.empty((nEl, nEl, hs), dtype=_F32)
        hinter = np.empty((nEl, nEl, hs), dtype=_F32)
        L.check(self._lib.frmc_export_data(self._handle, int(grid), L.ptr(hintra, L.c_f32p), L.ptr(hinter, L.c_f32p)),
                "export_data")
        return hintra, hinter

    def export_total(self, model, staged=False):
        out = np.empty(self._models[model].experimental.shape[0], dtype=_F32)
        L.check(self._lib.frmc_export_total(self._handle, int(model), int(bool(staged)), L.ptr(out, L.c_f32p)),
                "export_total")
        return out

    def get_coords(self):
        out = np.empty((self.numberOfAtoms, 3), dtype=_F32)
        L.check(self._lib.frmc_store_get_coords(self._handle, L.ptr(out, L.c_f32p)), "get_coords")
        return out

    def set_coords(self, boxCoords, basis=None):
        coords = L.as_array(boxCoords, "boxCoords", _F32, 2)
        b = None if basis is None else L.as_array(basis, "basis", _F32, 2)
        L.check(self._lib.frmc_store_set_coords(self._handle, L.ptr(coords, L.c_f32p), L.ptr(b, L.c_f32p)), "set_coords")

    # ------------------------------------------------------------------ timing (bench.py)
    TIMERS = {"delta": 0, "full": 1, "epilogue": 2, "commit": 3}

    def set_timing(self, on=True):
        L.check(self._lib.frmc_store_set_timing(self._handle, int(bool(on))), "set_timing")

    def get_timing(self, which):
        """(accumulated device ms, timed launches) of one kernel class since set_timing(True)."""
        ms = ctypes.c_double(0.0)
        n = ctypes.c_uint64(0)
        L.check(self._lib.frmc_store_get_timing(self._handle, self.TIMERS[which], ctypes.byref(ms), ctypes.byref(n)),
                "get_timing")
        return float(ms.value), int(n.value)

    @property
    def edge_overflow(self):
        return int(self._lib.frmc_store_edge_overflow(self._handle))

    @property
    def swept_pairs(self):
        """distance evaluations of the last compute_data (block culling skips the rest of n(n-1)/2)"""
        return int(self._lib.frmc_store_swept_pairs(self._handle))
