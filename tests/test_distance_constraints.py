"""Golden trajectories of the UNMODIFIED reference InterMolecularDistanceConstraint / IntraMolecularDistanceConstraint
(tests/gen_golden_distance_constraints.py, THF and SiOx example inputs) replayed through the device-backed mirror
(fullrmc_b200.constraints_distance): on the CPU with the oracle kernels (pins the mirror's host arithmetic), on the
GPU with the CUDA kernels.  Bar: data arrays, standard errors and rejection flags identical at every step."""
import os

import numpy as np
import pytest

NAMES = ["thf_inter", "thf_intra", "siox_inter"]
F32 = np.float32


def _replay(g, kernels, store=None):
    from fullrmc_b200.constraints_distance import DeviceMolecularDistanceConstraint
    box = g["boxCoords"].copy()
    c = DeviceMolecularDistanceConstraint(box, g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["typesIndex"], int(g["numberOfTypes"]),
                                          g["lowerLimitArray"], g["upperLimitArray"], g["typePairsIndex"],
                                          interMolecular=bool(g["interMolecular"]), flexible=bool(g["flexible"]), kernels=kernels,
                                          store=store)
    data, err = c.compute_data()
    assert np.array_equal(data["number"], g["start_number"]) and np.array_equal(data["distanceSum"], g["start_distanceSum"])
    assert F32(err) == F32(g["start_stdErr"])
    for s in range(g["steps/idx"].shape[0]):
        k = int(g["steps/k"][s])
        idx = g["steps/idx"][s, :k].astype(np.int32)
        moved = np.ascontiguousarray(g["steps/moved"][s, :k])
        c.compute_before_move(idx, idx)
        c.compute_after_move(idx, idx, moved)
        assert F32(c.afterMoveStandardError) == F32(g["steps/stdErr_after"][s]), "step %d" % s
        assert c.should_step_get_rejected(c.afterMoveStandardError) == bool(g["steps/rejected"][s]), "step %d" % s
        if bool(g["steps/accepted"][s]):
            c.accept_move(idx, idx)
            box[idx] = moved                                          # the engine moves the atoms (Engine.py:3337-3338)
        else:
            c.reject_move(idx, idx)
        assert np.array_equal(c.data["number"], g["steps/number"][s]) and np.array_equal(c.data["distanceSum"], g["steps/distanceSum"][s])
    assert F32(c.standardError) == F32(g["final_stdErr"])
    assert np.array_equal(c._get_constraint_value(), g["final_value"])


@pytest.mark.parametrize("name", NAMES)
def test_mirror_with_oracle_kernels_reproduces_reference_classes(name, golden_dir, orc):
    _replay(np.load(os.path.join(golden_dir, "distance_constraints_%s.npz" % name)), orc)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_mirror_on_device_reproduces_reference_classes(name, golden_dir):
    from fullrmc_b200.Core import atomic_distances
    _replay(np.load(os.path.join(golden_dir, "distance_constraints_%s.npz" % name)), atomic_distances)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_mirror_on_the_device_store_reproduces_reference_classes(name, golden_dir):
    """the constraint registered on a device store (csrc/storedist.cu): a move's four quantities from ONE pass over the
    resident records, no coordinate upload; the engine's boxCoordinates array is never read in the loop"""
    from fullrmc_b200.Core import atomic_distances
    from fullrmc_b200.store import DeviceStore
    g = np.load(os.path.join(golden_dir, "distance_constraints_%s.npz" % name))
    nEl = int(g["elementIndex"].max()) + 1
    with DeviceStore(g["boxCoords"], g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["elementIndex"], nEl) as store:
        _replay(g, atomic_distances, store=store)
        # the store moved the accepted atoms itself
        final = g["boxCoords"].copy()
        for s in range(g["steps/idx"].shape[0]):
            if bool(g["steps/accepted"][s]):
                k = int(g["steps/k"][s])
                final[g["steps/idx"][s, :k]] = g["steps/moved"][s, :k]
        assert np.array_equal(store.get_coords(), final)


@pytest.mark.gpu
def test_store_pass_equals_stateless_kernels_next_to_histogram_constraints():
    """a store that ALSO carries a histogram model (moves committed by accept(), deferred to the next launch): the
    distance pass sees the committed coordinates, and its four quantities equal the stateless kernels' on the same
    configuration; group moves of 5 atoms, two molecule types, intra + inter flags"""
    import time
    from fullrmc_b200 import synthetic
    from fullrmc_b200.Core import atomic_distances as ad
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    basis = np.array([[52, 0, 0], [7, 50, 0], [-5, 9, 49]], dtype=F32)
    s = synthetic.random_system(20000, 11, basis, n_elements=3, molecule_size=5)
    grid = synthetic.RGrid(0.0, 0.05, 200)
    rng = np.random.default_rng(3)
    nT = 3
    lower = np.zeros((nT, nT, 1), F32)
    upper = (1.6 + 0.3 * rng.random((nT, nT, 1))).astype(F32)
    upper = ((upper + upper.transpose(1, 0, 2)) / 2).astype(F32)
    flags = dict(interMolecular=True, intraMolecular=True, reduceDistance=False, reduceDistanceToUpper=True,
                 reduceDistanceToLower=False, countWithinLimits=True)
    common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume,
                  rho0=s.numberDensity, shell_centers=grid.shellCenters, shell_volumes=grid.shellVolumes)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 3) as store:
        gi = store.add_grid(grid.minDistance, grid.maxDistance, grid.bin, grid.hs)
        store.add_model(gi, ModelSpec("PDF", experimental=np.zeros(grid.hs, F32), **common))
        store.compute_data()
        cid = store.distance_add(s.elementIndex, nT, lower, upper, **flags)
        box = s.boxCoords.copy()
        kw = dict(basis=s.basis, isPBC=True, numberOfElements=nT, lowerLimit=lower, upperLimit=upper, **flags)
        t_store = 0.0
        for step in range(40):
            m = int(rng.integers(0, 20000 // 5))
            idx = np.arange(5 * m, 5 * m + 5, dtype=np.int32)
            moved = (box[idx] + rng.normal(0, 0.01, (5, 3))).astype(F32)
            t0 = time.perf_counter()
            counts, sums = store.distance_move(cid, idx, moved)
            t_store += time.perf_counter() - t0
            after = box.copy(); after[idx] = moved
            for which, coords in ((0, box), (2, after)):
                ni, di, ne, de = ad.multiple_atomic_distances_coords(indexes=idx, boxCoords=coords, moleculeIndex=s.moleculeIndex,
                                                                     elementIndex=s.elementIndex, allAtoms=True, **kw)
                fi, fd, fe, fde = ad.full_atomic_distances_coords(boxCoords=np.ascontiguousarray(coords[idx]),
                                                                  moleculeIndex=np.ascontiguousarray(s.moleculeIndex[idx]),
                                                                  elementIndex=np.ascontiguousarray(s.elementIndex[idx]), **kw)
                assert np.array_equal(counts[which, 0], ni) and np.array_equal(counts[which, 1], ne), "step %d M counts" % step
                assert np.array_equal(sums[which, 0], di) and np.array_equal(sums[which, 1], de), "step %d M sums" % step
                assert np.array_equal(counts[which + 1, 0], fi) and np.array_equal(counts[which + 1, 1], fe), "step %d F counts" % step
                assert np.array_equal(sums[which + 1, 0], fd) and np.array_equal(sums[which + 1, 1], fde), "step %d F sums" % step
            # the histogram constraint tries the same move; every other one is accepted (the commit is deferred)
            store.propose(idx, moved)
            if step % 2 == 0:
                store.accept(); box = after
            else:
                store.reject()
        assert np.array_equal(store.get_coords(), box)
        print("store distance pass: %.1f us per move (20000 atoms, 5-atom groups)" % (1e6 * t_store / 40))


@pytest.mark.gpu
def test_store_pass_latency_at_cfg4_size():
    """one atom of 100 000 (cfg4): the whole call -- flush, sweep, group pairs, sort, ordered sums, results on the host --
    stays in the tens of microseconds (the stateless per-move call uploads the coordinates: ~1 ms)"""
    import time
    from fullrmc_b200 import synthetic
    from fullrmc_b200.store import DeviceStore
    s = synthetic.cfg4()
    nT = 5
    lower = np.zeros((nT, nT, 1), F32)
    upper = np.full((nT, nT, 1), 1.5, F32)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 5) as store:
        cid = store.distance_add(s.elementIndex, nT, lower, upper, interMolecular=True, intraMolecular=False,
                                 reduceDistanceToUpper=True)
        rng = np.random.default_rng(0)
        idx = rng.integers(0, s.numberOfAtoms, 300).astype(np.int32)
        moved = (s.boxCoords[idx] + rng.normal(0, 0.001, (300, 3))).astype(F32)
        for it in range(20):
            store.distance_move(cid, idx[it:it + 1], moved[it:it + 1])
        t0 = time.perf_counter()
        for it in range(20, 300):
            store.distance_move(cid, idx[it:it + 1], moved[it:it + 1])
        us = 1e6 * (time.perf_counter() - t0) / 280
        print("store distance pass at cfg4: %.1f us per move" % us)
        assert us < 200.0


@pytest.mark.gpu
def test_store_pass_after_atoms_were_removed():
    """after removals the engine's relative indexes address the remaining atoms and the removed records pair with
    nothing: the store pass equals the stateless kernels on the engine's np.delete'd arrays"""
    from fullrmc_b200 import synthetic
    from fullrmc_b200.Core import atomic_distances as ad
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    basis = np.array([[40, 0, 0], [6, 39, 0], [-4, 8, 38]], dtype=F32)
    s = synthetic.random_system(6000, 3, basis, n_elements=3, molecule_size=2)
    grid = synthetic.RGrid(0.0, 0.05, 160)
    nT = 3
    lower = np.zeros((nT, nT, 1), F32)
    upper = np.full((nT, nT, 1), 2.0, F32)
    flags = dict(interMolecular=True, intraMolecular=True, reduceDistance=False, reduceDistanceToUpper=True,
                 reduceDistanceToLower=False, countWithinLimits=True)
    common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume,
                  rho0=s.numberDensity, shell_centers=grid.shellCenters, shell_volumes=grid.shellVolumes)
    rng = np.random.default_rng(9)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 3) as store:
        gi = store.add_grid(grid.minDistance, grid.maxDistance, grid.bin, grid.hs)
        store.add_model(gi, ModelSpec("PDF", experimental=np.zeros(grid.hs, F32), **common))
        store.compute_data()
        cid = store.distance_add(s.elementIndex, nT, lower, upper, **flags)
        box, mol, el = s.boxCoords.copy(), s.moleculeIndex.copy(), s.elementIndex.copy()
        for victim in (4000, 17, 2999, 17):
            store.propose_amputation(victim); store.accept_amputation()
            box, mol, el = np.delete(box, victim, axis=0), np.delete(mol, victim), np.delete(el, victim)
        assert store.numberOfAtoms == box.shape[0] and np.array_equal(store.get_coords(), box)
        kw = dict(basis=s.basis, isPBC=True, numberOfElements=nT, lowerLimit=lower, upperLimit=upper, **flags)
        for step in range(12):
            idx = np.sort(rng.choice(box.shape[0], 3, replace=False)).astype(np.int32)
            moved = (box[idx] + rng.normal(0, 0.02, (3, 3))).astype(F32)
            counts, sums = store.distance_move(cid, idx, moved)
            after = box.copy(); after[idx] = moved
            for which, coords in ((0, box), (2, after)):
                ni, di, ne, de = ad.multiple_atomic_distances_coords(indexes=idx, boxCoords=coords, moleculeIndex=mol, elementIndex=el,
                                                                     allAtoms=True, **kw)
                assert np.array_equal(counts[which, 0], ni) and np.array_equal(counts[which, 1], ne), "step %d" % step
                assert np.array_equal(sums[which, 0], di) and np.array_equal(sums[which, 1], de), "step %d" % step
