"""Atom removal and persisted state: trajectories that mix moves with removals, recorded from
the UNMODIFIED reference classes and the reference Engine's own bookkeeping (tests/gen_golden_removal.py:
compute_as_if_amputated / accept_amputation / reject_amputation, Engine._on_collector_collect_atom), replayed

* on the CPU through the oracle restatement (pins how the row leaves the histograms and which constants --
  weighting scheme, D_ij, number density -- each evaluation uses), and
* on the GPU through the device store (frmc_propose_amputation / frmc_accept_amputation / frmc_model_set_constants),

chi^2 of every step, weighting schemes, final data arrays and totals bit for bit.  Also: the full histogram of a store
with removed atoms, rebuilding the store after removals, runs of proposals after removals, and resuming from saved
data["intra"] / data["inter"] (frmc_import_data).
"""
import os

import numpy as np
import pytest

from oracle import epilogue as ep
from test_golden_constraints import _Golden, _constraint_desc, _oracle_total, _system

F32 = np.float32
NAMES = ["niti", "niti_fit", "siox", "synth"]


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "removal_%s.npz" % name))
    return _Golden((k, z[k]) for k in z.files)


def _weights(g, ci, elements):
    return {e: float(w) for e, w in zip(elements, g["c%d/elementsWeight" % ci])}


@pytest.fixture
def spill_oracle(orc):
    orc.set_emulate_spill(True)
    yield orc
    orc.set_emulate_spill(False)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_restatement_reproduces_reference_removals(name, golden_dir, spill_oracle):
    orc = spill_oracle
    g = _load(golden_dir, name)
    elements, n_per = _system(g)
    volume, rho_engine = F32(g["volume"]), F32(g["numberDensity"])
    box = g["boxCoords"].copy()
    basis, pbc = g["basis"], bool(g["isPBC"])
    mol, el = g["moleculeIndex"].copy(), g["elementIndex"].copy()
    nc = int(g["n_constraints"])
    allow_fit = bool(g["allow_fit"])
    descs = [_constraint_desc(g, ci) for ci in range(nc)]
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)

    def full(d, coords, mol, el):
        return orc.full_pairs_histograms_coords(boxCoords=coords, basis=basis, isPBC=pbc, moleculeIndex=mol, elementIndex=el,
                                                numberOfElements=len(elements), minDistance=d["minDistance"],
                                                maxDistance=d["maxDistance"], bin=d["bin"], histSize=int(d["histSize"]),
                                                ncores=orc.max_threads())
    data = [list(full(d, box, mol, el)) for d in descs]
    sfs = [F32(d["scaleFactor"]) for d in descs]
    accepted = 0
    for s in range(g["steps/idx"].shape[0]):
        rel = int(g["steps/idx"][s])
        idx = np.array([rel], np.int32)
        removal = int(g["steps/kind"][s]) == 1
        staged, used, new_w = [], [], []
        if removal:
            counts = dict(n_per)
            counts[elements[int(el[rel])]] -= 1
            rho_amp = F32((box.shape[0] - 1) / volume)                 # PairDistributionConstraints.py:1198
        for ci, d in enumerate(descs):
            args = (basis, pbc, mol, el, len(elements), d["minDistance"], d["maxDistance"], d["bin"], int(d["histSize"]))
            bi, be = ep.move_delta(fns, idx, box, *args)
            if removal:
                ni, ne = data[ci][0] - bi, data[ci][1] - be             # :1181-1184
                w = {k: F32(v) for k, v in ep.normalized_weighting(counts, _weights(g, ci, elements)).items()}
                dd = dict(d, weighting=w)
                tot, sf_used = _oracle_total(dd, ni, ne, elements, counts, volume, rho_amp, sf=sfs[ci],
                                             accepted=accepted if allow_fit else None)
                new_w.append(w)
            else:
                tmp = box.copy(); tmp[idx] = g["steps/moved"][s]
                ai, ae = ep.move_delta(fns, idx, tmp, *args)
                ni, ne = data[ci][0] - bi + ai, data[ci][1] - be + ae
                tot, sf_used = _oracle_total(d, ni, ne, elements, n_per, volume, rho_engine, sf=sfs[ci], accepted=accepted)
            chi = ep.standard_error(d["experimental"], tot, d["dataWeights"])
            assert F32(chi) == F32(g["steps/chi2_after"][s, ci]), "step %d (%s) constraint %d" % (s, "removal" if removal else "move", ci)
            assert F32(sf_used) == F32(g["steps/scale_used"][s, ci]), "step %d constraint %d scale factor" % (s, ci)
            staged.append([ni, ne]); used.append(F32(sf_used))
        if bool(g["steps/accepted"][s]):
            data, sfs = staged, used
            accepted += 1
            if removal:
                for ci, d in enumerate(descs):
                    d["weighting"] = new_w[ci]
                n_per = counts
                box, mol, el = np.delete(box, rel, axis=0), np.delete(mol, rel), np.delete(el, rel)
                if pbc:
                    rho_engine = F32(box.shape[0]) / F32(volume)                                  # Engine.py:795-796
            else:
                box[idx] = g["steps/moved"][s]
        for ci, d in enumerate(descs):
            assert np.array_equal(np.array([d["weighting"][str(p)] for p in d["pairs"]], F32), g["steps/pair_w"][s, ci]), \
                "step %d constraint %d weighting scheme" % (s, ci)
        assert F32(rho_engine) == F32(g["steps/numberDensity"][s])
    for ci, d in enumerate(descs):
        assert np.array_equal(data[ci][0], d["final_intra"]) and np.array_equal(data[ci][1], d["final_inter"])
        tot, _ = _oracle_total(d, data[ci][0], data[ci][1], elements, n_per, volume, rho_engine, sf=sfs[ci], accepted=accepted)
        assert np.array_equal(tot, d["final_total"])
        # the running arrays differ from a fresh histogram only in how the ordered [a,b] / [b,a] cells split; the
        # symmetrised sums the totals are built from agree
        ri, re_ = full(d, box, mol, el)
        assert np.array_equal(ri, d["recomputed_intra"]) and np.array_equal(re_, d["recomputed_inter"])
        sym = lambda a: a + np.transpose(a, (1, 0, 2))
        assert np.array_equal(sym(data[ci][0] + data[ci][1]), sym(ri + re_))
    assert np.array_equal(box, g["final_boxCoords"])


def _device_setup(g, persistent=False):
    from fullrmc_b200.constraints import DeviceBackend, make_device_constraint
    elements, n_per = _system(g)
    backend = DeviceBackend(g["boxCoords"], g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["elementIndex"], elements,
                            n_per, g["volume"], g["numberDensity"], persistent=persistent)
    backend.allowFittingScaleFactor = bool(g["allow_fit"])
    cons = []
    for ci in range(int(g["n_constraints"])):
        d = _constraint_desc(g, ci)
        cons.append((d, make_device_constraint(backend, d["kind"], d["experimental"], d["minDistance"], d["maxDistance"], d["bin"],
                                               int(d["histSize"]), d["shellCenters"], d["shellVolumes"], d["weighting"],
                                               dataWeights=d["dataWeights"], shapeArray=d["shapeArray"],
                                               scaleFactor=float(d["scaleFactor"]),
                                               qValues=d.get("qValues") if d["kind"] in ("SQ", "RSQ") else None,
                                               adjustScaleFactor=d["adjust"], elementsWeight=_weights(g, ci, elements))))
    return backend, cons


def _replay(g, backend, cons, first=0, last=None):
    last = g["steps/idx"].shape[0] if last is None else last
    for s in range(first, last):
        rel = np.array([int(g["steps/idx"][s])], np.int32)
        removal = int(g["steps/kind"][s]) == 1
        accept = bool(g["steps/accepted"][s])
        if removal:
            for d, c in cons:
                c.compute_as_if_amputated(rel, rel)
            for ci, (d, c) in enumerate(cons):
                assert F32(c.amputationStandardError) == F32(g["steps/chi2_after"][s, ci]), "step %d (removal) constraint %d" % (s, ci)
                assert F32(c.fittedScaleFactor) == F32(g["steps/scale_used"][s, ci]), "step %d constraint %d scale factor" % (s, ci)
            for d, c in cons:
                (c.accept_amputation if accept else c.reject_amputation)(rel, rel)
            if accept:
                backend._on_collector_collect_atom(int(rel[0]))
        else:
            moved = np.ascontiguousarray(g["steps/moved"][s:s + 1])
            for d, c in cons:
                c.compute_before_move(rel, rel)
                c.compute_after_move(rel, rel, moved)
            for ci, (d, c) in enumerate(cons):
                assert F32(c.afterMoveStandardError) == F32(g["steps/chi2_after"][s, ci]), "step %d (move) constraint %d" % (s, ci)
                assert F32(c.fittedScaleFactor) == F32(g["steps/scale_used"][s, ci]), "step %d constraint %d scale factor" % (s, ci)
            for d, c in cons:
                (c.accept_move if accept else c.reject_move)(rel, rel)
        for ci, (d, c) in enumerate(cons):
            assert np.array_equal(np.array([c.weighting[str(p)] for p in d["pairs"]], F32), g["steps/pair_w"][s, ci])
        assert F32(backend.numberDensity) == F32(g["steps/numberDensity"][s])


@pytest.mark.gpu
@pytest.mark.parametrize("persistent", [False, True])
@pytest.mark.parametrize("name", NAMES)
def test_device_store_reproduces_reference_removals(name, persistent, golden_dir):
    import fullrmc_b200
    previous = fullrmc_b200.set_edge_spill(True)
    try:
        g = _load(golden_dir, name)
        backend, cons = _device_setup(g, persistent)
        for ci, (d, c) in enumerate(cons):
            c.compute_data()
            assert F32(c.standardError) == F32(g["start_stdErr"][ci])
        _replay(g, backend, cons)
        assert backend.numberOfAtoms == g["final_boxCoords"].shape[0]
        for ci, (d, c) in enumerate(cons):
            data = c.data
            assert np.array_equal(data["intra"], d["final_intra"]) and np.array_equal(data["inter"], d["final_inter"])
            assert F32(c.standardError) == F32(d["final_stdErr"])
            assert F32(c.scaleFactor) == F32(d["final_scaleFactor"])
            # the recorded final total is a fresh evaluation with the engine's final state (number density, accepted
            # count): the committed total of an accepted removal was formed with rho0 = (N - 1) / volume, which is not
            # the engine's number density in a non-periodic system (Engine.py:795-796)
            c.compute_data(update=False)
            assert np.array_equal(c.get_constraint_total(), d["final_total"])
        assert np.array_equal(backend.store.get_coords(), g["final_boxCoords"])
        # a full histogram of the store with its holes = the reference's compute_data on the remaining atoms
        for ci, (d, c) in enumerate(cons):
            data, _ = c.compute_data()
            assert np.array_equal(data["intra"], d["recomputed_intra"]) and np.array_equal(data["inter"], d["recomputed_inter"])
        # ... and so does a store re-laid-out from the remaining atoms (set_coords after removals compacts the tables)
        backend.store.set_coords(g["final_boxCoords"])
        backend._dirty = True
        for ci, (d, c) in enumerate(cons):
            data, _ = c.compute_data()
            assert np.array_equal(data["intra"], d["recomputed_intra"]) and np.array_equal(data["inter"], d["recomputed_inter"])
        assert np.array_equal(backend.store.get_coords(), g["final_boxCoords"])
        backend.close()
    finally:
        fullrmc_b200.set_edge_spill(previous)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["niti", "synth"])
def test_resume_from_saved_data(name, golden_dir):
    """Save data["intra"/"inter"] and the coordinates in the middle of a trajectory (what the repository holds,
    Core/Constraint.py:275-288), build a NEW store from them without a histogram pass (frmc_import_data), finish the
    trajectory there: identical to the uninterrupted run."""
    import fullrmc_b200
    from fullrmc_b200.constraints import DeviceBackend, make_device_constraint
    previous = fullrmc_b200.set_edge_spill(True)
    try:
        g = _load(golden_dir, name)
        backend, cons = _device_setup(g)
        for d, c in cons:
            c.compute_data()
        half = g["steps/idx"].shape[0] // 2
        _replay(g, backend, cons, 0, half)
        saved = [dict(c.data) for d, c in cons]
        coords = backend.store.get_coords()
        elements = list(backend.elements)
        b2 = DeviceBackend(coords, g["basis"], bool(g["isPBC"]), backend.moleculesIndex, backend.elementsIndex, elements,
                           backend.numberOfAtomsPerElement, g["volume"], backend.numberDensity)
        b2.allowFittingScaleFactor = backend.allowFittingScaleFactor
        b2.accepted = backend.accepted
        b2.store.set_accepted(backend.accepted)
        cons2 = []
        for ci, (d, c) in enumerate(cons):
            cons2.append((d, make_device_constraint(b2, d["kind"], d["experimental"], d["minDistance"], d["maxDistance"], d["bin"],
                                                    int(d["histSize"]), d["shellCenters"], d["shellVolumes"], c.weighting,
                                                    dataWeights=d["dataWeights"], shapeArray=d["shapeArray"],
                                                    scaleFactor=float(c.scaleFactor),
                                                    qValues=d.get("qValues") if d["kind"] in ("SQ", "RSQ") else None,
                                                    adjustScaleFactor=d["adjust"], elementsWeight=c.elementsWeight)))
        for (d, c2), sv in zip(cons2, saved):
            c2.set_data(sv)
        assert not b2._dirty                                   # every grid imported: no full-histogram pass was made
        assert b2.store.swept_pairs == 0
        for (d, c), (_, c2) in zip(cons, cons2):
            assert F32(c2.standardError) == F32(c.standardError)
        backend.close()
        _replay(g, b2, cons2, half, None)
        for ci, (d, c2) in enumerate(cons2):
            data = c2.data
            assert np.array_equal(data["intra"], d["final_intra"]) and np.array_equal(data["inter"], d["final_inter"])
            assert F32(c2.standardError) == F32(d["final_stdErr"])
        assert np.array_equal(b2.store.get_coords(), g["final_boxCoords"])
        b2.close()
    finally:
        fullrmc_b200.set_edge_spill(previous)


@pytest.mark.gpu
def test_runs_of_proposals_after_removals_use_relative_indexes(golden_dir):
    """after removals the engine's (relative) indexes address the remaining atoms: frmc_run_batch and frmc_step agree"""
    import fullrmc_b200
    g = _load(golden_dir, "synth")
    rng = np.random.default_rng(5)
    outs = []
    for mode in ("batch", "steps"):
        backend, cons = _device_setup(g)
        for d, c in cons:
            c.compute_data()
        _replay(g, backend, cons)
        st = backend.store
        n = st.numberOfAtoms
        rng = np.random.default_rng(5)
        idx = rng.integers(0, n, 40).astype(np.int32)
        moved = (st.get_coords()[idx] + rng.normal(0, 0.004, (40, 3))).astype(F32)
        rnd = rng.random(40).astype(F32)
        total = F32(np.sum(st.committed_chi2(), dtype=F32))
        if mode == "batch":
            out = st.run_batch(idx, moved, total, rnd, tolerance=0.2)
            chi, dec = out["chi2"], out["decisions"]
        else:
            chi, dec, used = [], [], 0
            for j in range(40):
                c2 = st.propose(idx[j:j + 1], moved[j:j + 1])
                nt = F32(np.sum(c2, dtype=F32))
                d = 1
                if nt > total:
                    d = 0 if rnd[used] > F32(0.2) else 2
                    used += 1
                (st.accept if d else st.reject)()
                if d:
                    total = nt
                chi.append(c2); dec.append(d)
            chi, dec = np.array(chi, F32), np.array(dec, np.int32)
        outs.append((chi, dec, st.get_coords(), [c.data for d, c in cons]))
        backend.close()
    (c1, d1, x1, h1), (c2, d2, x2, h2) = outs
    assert np.array_equal(d1, d2) and np.array_equal(c1, c2) and np.array_equal(x1, x2)
    for a, b in zip(h1, h2):
        assert np.array_equal(a["intra"], b["intra"]) and np.array_equal(a["inter"], b["inter"])


@pytest.mark.gpu
def test_amputation_argument_errors(golden_dir):
    g = _load(golden_dir, "synth")
    backend, cons = _device_setup(g)
    st = backend.store
    with pytest.raises(RuntimeError):
        st.accept_amputation()                                  # nothing staged
    for d, c in cons:
        c.compute_data()
    with pytest.raises(ValueError):
        st.propose_amputation(st.numberOfAtoms)                 # index out of range
    st.propose_amputation(3)
    with pytest.raises(RuntimeError):
        st.propose(np.array([1], np.int32), np.zeros((1, 3), F32))   # an amputation is staged
    st.reject_amputation()
    with pytest.raises(ValueError):
        st.import_data(0, np.full((4, 4, 300), 0.5, F32), np.zeros((4, 4, 300), F32))   # not integer counts
    backend.close()
