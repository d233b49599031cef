"""The coordination-number pre-filter on the device store (csrc/storecoord.cu, DeviceStore.coordination_move; SURVEY.md
section 8f rank 3): the before / after counts of a move from ONE launch over the resident records.

* golden trajectories of the UNMODIFIED reference class (SiOx non-periodic, NiTi periodic) replay through the store mode
  of the mirror, data and standard error exactly;
* group moves next to a histogram model (deferred commits), atoms that are core and shell member of several definitions,
  a shell starting at 0 (the atom counts itself), non-finite bounds: equal to the stateless kernels (which are pinned to
  the compiled reference) on the same configuration, step by step;
* a store that is laid out again (set_coords) keeps answering correctly -- for the distance windows as well."""
import os
import time

import numpy as np
import pytest

from gen_golden_atomic_coordination import unpack_lists

F32 = np.float32
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ("siox", "niti"))
def test_store_mode_replays_reference_trajectory(case, golden_dir):
    from fullrmc_b200.constraints_coordination import DeviceAtomicCoordinationNumberConstraint
    from fullrmc_b200.store import DeviceStore
    g = np.load(os.path.join(golden_dir, "coordination_constraint_%s.npz" % case))
    box = g["boxCoords"].copy()
    n = box.shape[0]
    # the counts do not depend on elements or molecules: any labelling lays the store out
    el = (np.arange(n) % 2).astype(np.int32)
    mol = np.arange(n, dtype=np.int32)
    with DeviceStore(box, g["basis"], bool(g["isPBC"]), mol, el, 2) as store:
        c = DeviceAtomicCoordinationNumberConstraint(box, g["basis"], bool(g["isPBC"]), unpack_lists(g, "cores"), unpack_lists(g, "shells"),
                                                     g["lowerShells"], g["upperShells"], g["minAtoms"], g["maxAtoms"], g["weights"], store=store)
        data, err = c.compute_data()
        assert np.array_equal(data, g["start_data"]) and F32(err) == g["start_stdErr"]
        for step in range(len(g["steps/idx"])):
            idx = g["steps/idx"][step:step + 1].astype(np.int32)
            moved = g["steps/moved"][step:step + 1]
            c.compute_before_move(realIndexes=idx, relativeIndexes=idx)
            c.compute_after_move(realIndexes=idx, relativeIndexes=idx, movedBoxCoordinates=moved)
            assert F32(c.afterMoveStandardError) == g["steps/stdErr_after"][step], (case, step)
            if g["steps/accepted"][step]:
                c.accept_move(realIndexes=idx, relativeIndexes=idx)
                box[idx] = moved
            else:
                c.reject_move(realIndexes=idx, relativeIndexes=idx)
            assert np.array_equal(c.data, g["steps/data"][step]), (case, step)
        assert F32(c.standardError) == g["final_stdErr"]
        assert np.array_equal(store.get_coords(), box)                 # the store moved the accepted atoms itself
        recount, _ = c.compute_data(update=False)
        assert np.array_equal(recount, g["final_recount"])


def _definitions(s, rng):
    """five definitions over a 3-element system: element shells, overlapping explicit lists, a shell from 0, an unbounded one"""
    n = s.numberOfAtoms
    by_el = [np.flatnonzero(s.elementIndex == e).astype(np.int32) for e in range(3)]
    some = np.sort(rng.choice(n, n // 3, replace=False)).astype(np.int32)
    other = np.sort(rng.choice(n, n // 2, replace=False)).astype(np.int32)
    cores = [by_el[0], by_el[1], some, np.arange(n, dtype=np.int32), by_el[2]]
    shells = [by_el[1], by_el[1], other, some, np.arange(n, dtype=np.int32)]
    lower = [F32(1.0), F32(0.0), F32(0.5), F32(0.0), F32(2.0)]
    upper = [F32(3.0), F32(2.5), F32(3.5), F32(1.75), F32(np.inf)]
    return cores, shells, lower, upper


def _stateless_counts(idx, coords, s, cores, shells, lower, upper):
    from fullrmc_b200.Core import atomic_coordination as ac
    from fullrmc_b200.constraints_coordination import _membership
    out = np.zeros(len(cores), F32)
    ac.multi_atoms_coord_number_coords(indexes=idx, boxCoords=coords, basis=s.basis, isPBC=True, coresIndexes=cores, shellsIndexes=shells,
                                       lowerShells=lower, upperShells=upper, asCoreDefIdxs=_membership(cores, s.numberOfAtoms),
                                       inShellDefIdxs=_membership(shells, s.numberOfAtoms), coordNumData=out)
    return out


def test_store_pass_equals_stateless_kernels_next_to_histogram_constraints():
    from fullrmc_b200 import synthetic
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    basis = np.array([[42, 0, 0], [6, 40, 0], [-4, 7, 39]], dtype=F32)
    s = synthetic.random_system(6000, 5, basis, n_elements=3, molecule_size=5)
    grid = synthetic.RGrid(0.0, 0.05, 200)
    rng = np.random.default_rng(9)
    cores, shells, lower, upper = _definitions(s, rng)
    common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume,
                  rho0=s.numberDensity, shell_centers=grid.shellCenters, shell_volumes=grid.shellVolumes)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 3) as store:
        gi = store.add_grid(grid.minDistance, grid.maxDistance, grid.bin, grid.hs)
        store.add_model(gi, ModelSpec("PDF", experimental=np.zeros(grid.hs, F32), **common))
        store.compute_data()
        cid = store.coordination_add(cores, shells, lower, upper)
        box = s.boxCoords.copy()
        for step in range(24):
            k = (1, 5, 13)[step % 3]
            first = int(rng.integers(0, 6000 - k))
            idx = np.arange(first, first + k, dtype=np.int32)
            moved = (box[idx] + rng.normal(0, 0.02, (k, 3))).astype(F32)
            counts = store.coordination_move(cid, idx, moved).copy()
            after = box.copy(); after[idx] = moved
            assert np.array_equal(counts[0].astype(F32), _stateless_counts(idx, box, s, cores, shells, lower, upper)), "before, step %d" % step
            assert np.array_equal(counts[1].astype(F32), _stateless_counts(idx, after, s, cores, shells, lower, upper)), "after, step %d" % step
            assert counts[0].sum() > 0
            store.propose(idx, moved)                                  # every other move is accepted (the commit is deferred)
            if step % 2 == 0:
                store.accept(); box = after
            else:
                store.reject()
        assert np.array_equal(store.get_coords(), box)
        # the store laid out again (new coordinates, the records change places): masks and types follow the records
        perm_box = ((box + rng.normal(0, 0.3, box.shape)) % 1.0).astype(F32)
        store.set_coords(perm_box)
        idx = np.arange(100, 105, dtype=np.int32)
        moved = (perm_box[idx] + rng.normal(0, 0.02, (5, 3))).astype(F32)
        counts = store.coordination_move(cid, idx, moved).copy()
        after = perm_box.copy(); after[idx] = moved
        assert np.array_equal(counts[0].astype(F32), _stateless_counts(idx, perm_box, s, cores, shells, lower, upper))
        assert np.array_equal(counts[1].astype(F32), _stateless_counts(idx, after, s, cores, shells, lower, upper))


def test_distance_windows_follow_a_new_layout():
    """frmc_store_set_coords lays the store out again: the per-position type table of a registered distance constraint
    is rebuilt (csrc/storedist.cu) -- its four quantities still equal the stateless kernels'"""
    from fullrmc_b200 import synthetic
    from fullrmc_b200.Core import atomic_distances as ad
    from fullrmc_b200.store import DeviceStore
    basis = np.array([[40, 0, 0], [0, 41, 0], [0, 0, 42]], dtype=F32)
    s = synthetic.random_system(5000, 21, basis, n_elements=3, molecule_size=5)
    nT = 3
    lower = np.zeros((nT, nT, 1), F32)
    upper = np.full((nT, nT, 1), 1.8, F32)
    flags = dict(interMolecular=True, intraMolecular=True, reduceDistance=False, reduceDistanceToUpper=True,
                 reduceDistanceToLower=False, countWithinLimits=True)
    rng = np.random.default_rng(4)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 3) as store:
        cid = store.distance_add(s.elementIndex, nT, lower, upper, **flags)
        box = ((s.boxCoords + rng.normal(0, 0.3, s.boxCoords.shape)) % 1.0).astype(F32)
        store.set_coords(box)
        kw = dict(basis=s.basis, isPBC=True, numberOfElements=nT, lowerLimit=lower, upperLimit=upper, **flags)
        seen = 0
        for step in range(6):
            m = int(rng.integers(0, 1000))
            idx = np.arange(5 * m, 5 * m + 5, dtype=np.int32)
            moved = (box[idx] + rng.normal(0, 0.01, (5, 3))).astype(F32)
            counts, sums = store.distance_move(cid, idx, moved)
            ni, di, ne, de = ad.multiple_atomic_distances_coords(indexes=idx, boxCoords=box, moleculeIndex=s.moleculeIndex,
                                                                 elementIndex=s.elementIndex, allAtoms=True, **kw)
            assert np.array_equal(counts[0, 0], ni) and np.array_equal(counts[0, 1], ne)
            assert np.array_equal(sums[0, 0], di) and np.array_equal(sums[0, 1], de)
            seen += int(ni.sum() + ne.sum())
        assert seen > 0


def test_store_pass_latency_at_cfg4_size():
    """one atom of 100 000 (cfg4), three element-shell definitions: the whole call stays in the tens of microseconds (the
    stateless per-move call flattens the lists and uploads the coordinates: ~1 ms)"""
    from fullrmc_b200 import synthetic
    from fullrmc_b200.store import DeviceStore
    s = synthetic.cfg4()
    by_el = [np.flatnonzero(s.elementIndex == e).astype(np.int32) for e in range(5)]
    cores, shells = [by_el[0], by_el[1], by_el[2]], [by_el[1], by_el[2], by_el[0]]
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 5) as store:
        cid = store.coordination_add(cores, shells, [1.5] * 3, [3.5] * 3)
        rng = np.random.default_rng(0)
        idx = rng.integers(0, s.numberOfAtoms, 300).astype(np.int32)
        moved = (s.boxCoords[idx] + rng.normal(0, 0.001, (300, 3))).astype(F32)
        for it in range(20):
            store.coordination_move(cid, idx[it:it + 1], moved[it:it + 1])
        t0 = time.perf_counter()
        for it in range(20, 300):
            store.coordination_move(cid, idx[it:it + 1], moved[it:it + 1])
        us = 1e6 * (time.perf_counter() - t0) / 280
        print("store coordination pass at cfg4: %.1f us per move" % us)
        assert us < 200.0


def test_argument_errors():
    from fullrmc_b200 import synthetic
    from fullrmc_b200._lib import FullrmcB200Error
    from fullrmc_b200.store import DeviceStore
    basis = np.eye(3, dtype=F32) * 20
    s = synthetic.random_system(500, 1, basis, n_elements=2, molecule_size=1)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 2) as store:
        with pytest.raises(ValueError):
            store.coordination_add([np.array([1, 1], np.int32)], [np.array([2], np.int32)], [0.0], [1.0])      # an atom twice
        with pytest.raises(ValueError):
            store.coordination_add([np.array([600], np.int32)], [np.array([2], np.int32)], [0.0], [1.0])       # out of range
        with pytest.raises((ValueError, RuntimeError, FullrmcB200Error)):
            store.coordination_add([np.array([1], np.int32)] * 33, [np.array([2], np.int32)] * 33, [0.0] * 33, [1.0] * 33)
        cid = store.coordination_add([np.array([1], np.int32)], [np.array([2], np.int32)], [0.0], [1.0])
        with pytest.raises(ValueError):
            store.coordination_move(cid + 1, np.array([1], np.int32), s.boxCoords[1:2])


def test_store_pass_after_atoms_were_removed():
    """definitions registered on the full system; after removals the engine's relative indexes address the remaining atoms
    and the removed records take part in nothing: the store pass equals the stateless kernels on the engine's np.delete'd
    arrays with the lists re-numbered the way the engine's collector re-numbers them"""
    from fullrmc_b200 import synthetic
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    basis = np.array([[40, 0, 0], [6, 39, 0], [-4, 8, 38]], dtype=F32)
    s = synthetic.random_system(6000, 3, basis, n_elements=3, molecule_size=2)
    grid = synthetic.RGrid(0.0, 0.05, 160)
    rng = np.random.default_rng(13)
    cores, shells, lower, upper = _definitions(s, rng)
    common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume,
                  rho0=s.numberDensity, shell_centers=grid.shellCenters, shell_volumes=grid.shellVolumes)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 3) as store:
        gi = store.add_grid(grid.minDistance, grid.maxDistance, grid.bin, grid.hs)
        store.add_model(gi, ModelSpec("PDF", experimental=np.zeros(grid.hs, F32), **common))
        store.compute_data()
        cid = store.coordination_add(cores, shells, lower, upper)
        box = s.boxCoords.copy()
        alive = np.arange(6000)                                       # original index of every remaining atom
        for victim in (4000, 17, 2999, 17):
            store.propose_amputation(victim); store.accept_amputation()
            box, alive = np.delete(box, victim, axis=0), np.delete(alive, victim)
        assert store.numberOfAtoms == box.shape[0]
        new_index = np.full(6000, -1, np.int64); new_index[alive] = np.arange(alive.shape[0])
        renum = lambda lists: [np.ascontiguousarray(new_index[a][new_index[a] >= 0], dtype=np.int32) for a in lists]
        cores2, shells2 = renum(cores), renum(shells)

        class _S(object):
            basis, numberOfAtoms = s.basis, box.shape[0]

        for step in range(10):
            idx = np.sort(rng.choice(box.shape[0], 3, replace=False)).astype(np.int32)
            moved = (box[idx] + rng.normal(0, 0.02, (3, 3))).astype(F32)
            counts = store.coordination_move(cid, idx, moved).copy()
            after = box.copy(); after[idx] = moved
            assert np.array_equal(counts[0].astype(F32), _stateless_counts(idx, box, _S, cores2, shells2, lower, upper)), "before, step %d" % step
            assert np.array_equal(counts[1].astype(F32), _stateless_counts(idx, after, _S, cores2, shells2, lower, upper)), "after, step %d" % step
