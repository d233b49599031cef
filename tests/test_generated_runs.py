"""Device-generated runs of moves: group selection, translation, transform_coordinates,
evaluation, decision and move application on the device under the counter-based random-number contract of
fullrmc_b200/rng.py.

The fixtures (tests/gen_golden_generated.py) are whole `Engine.run`s of the UNMODIFIED reference engine equipped with the
plug-ins of fullrmc_b200/engine_plugins.py.  They are replayed

* on the CPU with the numpy statement of the contract (rng.generate_step) + the oracle's histograms and totals + the
  engine's rule -- pins rng.py to what the reference Engine did with the plug-ins, and
* on the GPU through `DeviceStore.run_generated` (one call, and cut into several calls): accepted count, standard
  errors, data arrays, box AND real coordinates bit for bit.
"""
import os

import numpy as np
import pytest

from oracle import epilogue as ep
from fullrmc_b200 import rng
from test_golden_constraints import _Golden, _constraint_desc, _oracle_total, _system

F32 = np.float32
NAMES = ["niti", "niti_sf", "thf", "siox", "synth"]


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "generated_%s.npz" % name))
    return _Golden((k, z[k]) for k in z.files)


def _amplitude(g):
    a = np.atleast_1d(g["amplitude"]).astype(F32)
    return (F32(0.0), a[0]) if a.shape[0] == 1 else (a[0], a[1])


def test_philox_known_answers():
    """Random123's known-answer vectors for philox4x32-10"""
    f = lambda t: " ".join("%08x" % x for x in t)
    assert f(rng.philox4x32((0, 0, 0, 0), (0, 0))) == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert f(rng.philox4x32((0xffffffff,) * 4, (0xffffffff,) * 2)) == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert f(rng.philox4x32((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))) == \
        "d16cfe09 94fdcceb 5001e420 24126ea1"


def test_transform_coordinates_restatement_equals_the_compiled_reference():
    from oracle import build_ref
    if not build_ref.is_built():
        build_ref.build()
    import sys
    if build_ref.OUT not in sys.path:
        sys.path.insert(0, build_ref.OUT)
    try:
        from fullrmc.Core import boundary_conditions_collection as bcc
    except Exception:
        pytest.skip("compiled reference boundary_conditions_collection not available")
    r = np.random.default_rng(0)
    for _ in range(5):
        m = r.normal(0, 0.05, (3, 3)).astype(F32)
        c = r.normal(0, 60, (257, 3)).astype(F32)
        assert np.array_equal(rng.transform_coordinates(m, c), bcc.transform_coordinates(m, c))


def test_translation_vector_is_within_the_amplitude_range():
    for c in range(200):
        w = rng.step_words(12345, c)
        v = rng.translation_vector(w, 0.05, 0.25)
        nrm = float(np.linalg.norm(v.astype(np.float64)))
        assert 0.05 - 1e-6 <= nrm < 0.25 + 1e-6
        assert 0.0 <= float(rng.acceptance_number(w)) < 1.0


@pytest.fixture
def spill_oracle(orc):
    orc.set_emulate_spill(True)
    yield orc
    orc.set_emulate_spill(False)


@pytest.mark.parametrize("name", ["siox", "synth"])
def test_numpy_statement_of_the_contract_reproduces_the_reference_engine(name, golden_dir, spill_oracle):
    orc = spill_oracle
    g = _load(golden_dir, name)
    elements, n_per = _system(g)
    volume, rho0 = F32(g["volume"]), F32(g["numberDensity"])
    basis, pbc, mol, el = g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["elementIndex"]
    box, real = g["boxCoords"].copy(), g["realCoords"].copy()
    rb = g["reciprocalBasis"] if pbc else None
    descs = [_constraint_desc(g, ci) for ci in range(int(g["n_constraints"]))]
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)
    data, chis = [], []
    for d in descs:
        hi, he = orc.full_pairs_histograms_coords(boxCoords=box, basis=basis, isPBC=pbc, moleculeIndex=mol, elementIndex=el,
                                                  numberOfElements=len(elements), minDistance=d["minDistance"], maxDistance=d["maxDistance"],
                                                  bin=d["bin"], histSize=int(d["histSize"]), ncores=orc.max_threads())
        data.append([hi, he])
        tot, _ = _oracle_total(d, hi, he, elements, n_per, volume, rho0)
        chis.append(F32(ep.standard_error(d["experimental"], tot, d["dataWeights"])))
    total = F32(np.sum(np.array(chis, F32)))                       # Engine.compute_total_standard_error, varianceSquared = 1
    lo, hi_amp = _amplitude(g)
    seed, c0 = int(g["seed"]), int(g["first_counter"])
    accepted = 0
    for s in range(int(g["n_steps"])):
        grp, idx, mreal, mbox, u = rng.generate_step(seed, c0 + s, g["group_offsets"], g["group_indexes"], real, rb, lo, hi_amp)
        tmp = box.copy(); tmp[idx] = mbox
        staged, new_chis = [], []
        for ci, d in enumerate(descs):
            args = (basis, pbc, mol, el, len(elements), d["minDistance"], d["maxDistance"], d["bin"], int(d["histSize"]))
            bi, be = ep.move_delta(fns, idx, box, *args)
            ai, ae = ep.move_delta(fns, idx, tmp, *args)
            ni, ne = data[ci][0] - bi + ai, data[ci][1] - be + ae
            tot, _ = _oracle_total(d, ni, ne, elements, n_per, volume, rho0)
            staged.append([ni, ne]); new_chis.append(F32(ep.standard_error(d["experimental"], tot, d["dataWeights"])))
        new_total = F32(np.sum(np.array(new_chis, F32)))
        accept = True
        if new_total > total:
            accept = not (float(u) > 0.0)                             # Engine.py:3310-3315, tolerance 0
        if accept:
            data, chis, total, box = staged, new_chis, new_total, tmp
            real[idx] = mreal
            accepted += 1
    assert accepted == int(g["accepted"])
    assert np.array_equal(box, g["final_boxCoords"]) and np.array_equal(real, g["final_realCoords"])
    for ci, d in enumerate(descs):
        assert np.array_equal(data[ci][0], d["final_intra"]) and np.array_equal(data[ci][1], d["final_inter"])
        assert F32(chis[ci]) == F32(d["final_stdErr"])
    assert F32(total) == F32(g["totalStandardError"])


def _device_store(g):
    from fullrmc_b200.constraints import DeviceBackend, make_device_constraint
    elements, n_per = _system(g)
    backend = DeviceBackend(g["boxCoords"], g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["elementIndex"], elements,
                            n_per, g["volume"], g["numberDensity"])
    cons = []
    for ci in range(int(g["n_constraints"])):
        d = _constraint_desc(g, ci)
        cons.append((d, make_device_constraint(backend, d["kind"], d["experimental"], d["minDistance"], d["maxDistance"], d["bin"],
                                               int(d["histSize"]), d["shellCenters"], d["shellVolumes"], d["weighting"],
                                               dataWeights=d["dataWeights"], shapeArray=d["shapeArray"], scaleFactor=float(d["scaleFactor"]),
                                               qValues=d.get("qValues") if d["kind"] in ("SQ", "RSQ") else None,
                                               adjustScaleFactor=d["adjust"])))
    st = backend.store
    off = g["group_offsets"]
    st.set_groups([g["group_indexes"][off[i]:off[i + 1]] for i in range(off.shape[0] - 1)])
    if bool(g["isPBC"]):
        st.set_real_coords(np.ascontiguousarray(g["realCoords"]), np.ascontiguousarray(g["reciprocalBasis"]))
    else:
        st.set_real_coords()
    return backend, cons


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [1, 7])
@pytest.mark.parametrize("name", NAMES)
def test_device_generated_run_reproduces_the_reference_engine(name, chunks, golden_dir):
    import fullrmc_b200
    previous = fullrmc_b200.set_edge_spill(True)
    try:
        g = _load(golden_dir, name)
        backend, cons = _device_store(g)
        st = backend.store
        chi0 = backend._compute_data()
        total = F32(np.sum(chi0, dtype=F32))
        n, seed, c0 = int(g["n_steps"]), int(g["seed"]), int(g["first_counter"])
        amp = _amplitude(g)
        cuts = np.linspace(0, n, chunks + 1).astype(int)
        accepted, groups, rands = 0, [], []
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b == a:
                continue
            out = st.run_generated(int(b - a), seed, c0 + int(a), amp, total)
            total = out["total"]
            accepted += int((out["decisions"] > 0).sum())
            groups.append(out["groups"]); rands.append(out["rand"])
        groups, rands = np.concatenate(groups), np.concatenate(rands)
        # the numbers the device drew are the contract's
        ng = g["group_offsets"].shape[0] - 1
        for s in (0, 1, n // 2, n - 1):
            w = rng.step_words(seed, c0 + s)
            assert int(groups[s]) == rng.group_index(w[0], ng) and F32(rands[s]) == rng.acceptance_number(w)
        assert accepted == int(g["accepted"])
        assert F32(total) == F32(g["totalStandardError"])
        assert np.array_equal(st.get_coords(), g["final_boxCoords"])
        assert np.array_equal(st.get_real_coords(), g["final_realCoords"])
        chi = st.committed_chi2()
        for ci, (d, c) in enumerate(cons):
            data = c.data
            assert np.array_equal(data["intra"], d["final_intra"]) and np.array_equal(data["inter"], d["final_inter"])
            assert F32(chi[ci]) == F32(d["final_stdErr"])
            assert F32(st.get_scale(c._model)[0]) == F32(d["final_scaleFactor"])
        backend.close()
    finally:
        fullrmc_b200.set_edge_spill(previous)


@pytest.mark.gpu
def test_generated_run_needs_groups_and_real_coordinates(golden_dir):
    g = _load(golden_dir, "synth")
    from fullrmc_b200.constraints import DeviceBackend
    backend, cons = _device_store(g)
    st = backend.store
    total = F32(np.sum(backend._compute_data(), dtype=F32))
    # a move accepted through another entry point invalidates the real coordinates
    idx = np.array([5], np.int32)
    st.propose(idx, g["boxCoords"][idx] + F32(0.001)); st.accept()
    with pytest.raises(RuntimeError):
        st.run_generated(8, 1, 0, 0.2, total)
    st.set_real_coords(np.ascontiguousarray(g["realCoords"]), np.ascontiguousarray(g["reciprocalBasis"]))
    out = st.run_generated(8, 1, 0, 0.2, total)
    assert out["decisions"].shape == (8,)
    with pytest.raises(ValueError):
        st.run_generated(8, 1, 0, (0.3, 0.1), total)              # empty amplitude range
    backend.close()


@pytest.mark.gpu
def test_stateless_transform_coordinates(golden_dir):
    from fullrmc_b200.Core import boundary_conditions_collection as bcc
    r = np.random.default_rng(1)
    m = r.normal(0, 0.05, (3, 3)).astype(F32)
    c = r.normal(0, 60, (10001, 3)).astype(F32)
    assert np.array_equal(bcc.transform_coordinates(m, c), rng.transform_coordinates(m, c))
    assert bcc.transform_coordinates(m, np.zeros((0, 3), F32)).shape == (0, 3)
    with pytest.raises(ValueError):
        bcc.transform_coordinates(m.astype(np.float64), c)
    with pytest.raises(TypeError):
        bcc.transform_coordinates(None, c)
