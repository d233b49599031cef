"""GPU parity tests of the stateful device path (DeviceStore: compute_data, propose, accept,
reject) against the reference sequence compute_before_move / compute_after_move / accept_move
(PairDistributionConstraints.py:1044-1152) restated with the oracle.

Bars: running histograms bit-exact (ordered arrays, which implies the symmetrised sums the
physics uses); G(r), g(r), S(Q) totals and chi^2 bit-exact against the numpy restatement
(the north_star tolerance is 1e-6 relative; the device epilogue mirrors numpy's fp32 operation
and summation order, so equality is asserted and the tolerance is the fallback bar)."""
import numpy as np
import pytest

import cases as C
from oracle import epilogue as ep

pytestmark = pytest.mark.gpu

F32 = np.float32
ELEMENTS = ["O", "Si", "Ti", "Ni", "Zr"]
CASES = {c["name"]: c for c in C.make_cases()}


def _system_meta(case):
    nEl = case["numberOfElements"]
    els = ELEMENTS[:nEl]
    counts = np.bincount(case["elementIndex"], minlength=nEl)
    n_per = {els[i]: int(max(counts[i], 2)) for i in range(nEl)}      # empty classes: keep D_ij finite
    weights = {"O": 8.0, "Si": 14.0, "Ti": 22.0, "Ni": 28.0, "Zr": 40.0}
    wdict = ep.normalized_weighting(n_per, {e: weights[e] for e in els})
    wdict = {k: F32(v) for k, v in wdict.items()}
    if case["isPBC"]:
        volume = F32(abs(np.linalg.det(case["basis"].astype(np.float64))))
    else:
        volume = F32(case["boxCoords"].shape[0] / 0.0333679)
    rho0 = F32(case["boxCoords"].shape[0]) / F32(volume)
    return els, n_per, wdict, volume, rho0


def _grid_arrays(case):
    hs = case["histSize"]
    edges = (case["minDistance"] + case["bin"] * np.arange(hs + 1, dtype=np.float64)).astype(F32)
    centers = ((edges[:-1] + edges[1:]) / F32(2.)).astype(F32)
    return centers, ep.shell_arrays_from_edges(edges)


PERSISTENT = False       # flipped by test_persistent_kernel_*: every store built below keeps one kernel resident


def _build(case, kinds, rng, with_weights=False, with_shape=False, scale=1.0):
    """DeviceStore + ModelSpecs + matching oracle closures for the requested model kinds"""
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    els, n_per, wdict, volume, rho0 = _system_meta(case)
    centers, shellv = _grid_arrays(case)
    hs = case["histSize"]
    store = DeviceStore(case["boxCoords"], case["basis"], case["isPBC"], case["moleculeIndex"], case["elementIndex"],
                        case["numberOfElements"])
    if PERSISTENT:
        store.set_persistent(True)
    g = store.add_grid(case["minDistance"], case["maxDistance"], case["bin"], hs)
    q = np.linspace(0.5, 20.0, 97).astype(F32)
    gr2sq = ep.gr2sq_matrix(q, centers)
    oracles = []
    for kind in kinds:
        n_out = hs if kind in ("PDF", "PCF") else q.shape[0]
        exp = rng.normal(0.0 if kind != "PCF" else 1.0, 0.5, n_out).astype(F32)
        dw = rng.random(n_out).astype(F32) if with_weights else None
        shape = (0.01 * rng.standard_normal(hs)).astype(F32) if (with_shape and kind in ("PDF", "PCF")) else None
        spec = ModelSpec(kind, els, n_per, wdict, volume, rho0, centers, shellv, exp, data_weights=dw,
                         shape_array=shape, scale_factor=scale, q_values=q if kind in ("SQ", "RSQ") else None)
        store.add_model(g, spec)
        common = dict(elements=els, n_per_element=n_per, weighting=wdict, volume=volume, rho0=rho0,
                      shell_centers=centers, shell_volumes=shellv)

        def total(intra, inter, kind=kind, shape=shape, common=common):
            if kind == "PDF":
                return ep.total_Gr(intra, inter, shape_array=shape, scale_factor=scale, **common)
            if kind == "PCF":
                return ep.total_gr(intra, inter, shape_array=shape, scale_factor=scale, **common)
            return ep.total_Sq(intra, inter, gr2sq=gr2sq, scale_factor=scale, reduced=(kind == "RSQ"), **common)
        oracles.append((total, exp, dw))
    return store, oracles


def _check_models(store, oracles, intra, inter, chi2, staged):
    for m, (total, exp, dw) in enumerate(oracles):
        want = total(intra, inter)
        got = store.export_total(m, staged=staged)
        scale = max(1e-30, float(np.max(np.abs(want))))
        assert np.max(np.abs(got - want)) <= 1e-6 * scale, "model %d total outside 1e-6" % m
        assert np.array_equal(got, want), "model %d total not bit-exact (max diff %g)" % (m, np.max(np.abs(got - want)))
        want_chi = ep.standard_error(exp, want, dw)
        assert abs(float(chi2[m]) - float(want_chi)) <= 1e-6 * abs(float(want_chi))
        assert F32(chi2[m]) == F32(want_chi), "model %d chi2 %r != %r" % (m, chi2[m], want_chi)


def _hist_kw(case):
    return dict(basis=case["basis"], isPBC=case["isPBC"], numberOfElements=case["numberOfElements"],
                minDistance=case["minDistance"], maxDistance=case["maxDistance"], bin=case["bin"],
                histSize=case["histSize"])


def _run_sequence(case, kinds, orc, n_moves=12, seed=0, sigma=0.02, **model_kw):
    rng = np.random.default_rng(seed)
    store, oracles = _build(case, kinds, rng, **model_kw)
    kw = _hist_kw(case)
    mol, el = case["moleculeIndex"], case["elementIndex"]
    box = case["boxCoords"].copy()
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)

    chi2 = store.compute_data()
    data_i, data_e = orc.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=el, **kw)
    gi, ge = store.export_data(0)
    assert np.array_equal(gi, data_i) and np.array_equal(ge, data_e)
    _check_models(store, oracles, data_i, data_e, chi2, staged=False)

    accepted = 0
    for step in range(n_moves):
        idx = C.group_for(case, rng)
        moved = (box[idx] + rng.normal(0, sigma, (idx.shape[0], 3)).astype(F32)).astype(F32)
        # reference sequence (PairDistributionConstraints.py:1044-1129)
        bi, be = ep.move_delta(fns, idx, box, kw["basis"], kw["isPBC"], mol, el, kw["numberOfElements"],
                               kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
        tmp = box.copy(); tmp[idx] = moved
        ai, ae = ep.move_delta(fns, idx, tmp, kw["basis"], kw["isPBC"], mol, el, kw["numberOfElements"],
                               kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
        new_i = data_i - bi + ai
        new_e = data_e - be + ae
        chi2 = store.propose(idx, moved)
        _check_models(store, oracles, new_i, new_e, chi2, staged=True)
        if step % 3 != 2:                      # accept two moves out of three
            store.accept()
            data_i, data_e, box = new_i, new_e, tmp
            accepted += 1
        else:
            store.reject()
        gi, ge = store.export_data(0)
        assert np.array_equal(gi, data_i) and np.array_equal(ge, data_e), "running histograms diverged at step %d" % step
    assert accepted > 0
    assert np.array_equal(store.get_coords(), box)
    assert store.edge_overflow == 0
    # the incrementally updated state equals a from-scratch recomputation (symmetrised)
    fi, fe = orc.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=el, **kw)
    sym = lambda h: h + h.transpose(1, 0, 2)
    assert np.array_equal(sym(data_i), sym(fi)) and np.array_equal(sym(data_e), sym(fe))
    store.close()


def test_atomic_orthorhombic_pdf_and_sq(orc):
    _run_sequence(CASES["ortho_atomic"], ["PDF", "SQ"], orc, seed=1)


def test_molecular_triclinic_pcf_with_weights(orc):
    _run_sequence(CASES["tri_molecular"], ["PCF"], orc, seed=2, with_weights=True)


def test_unwrapped_triclinic_reduced_sq_scaled(orc):
    _run_sequence(CASES["tri_unwrapped"], ["RSQ", "PDF"], orc, seed=3, scale=0.93, with_shape=True)


def test_unwrapped_orthorhombic(orc):
    _run_sequence(CASES["ortho_unwrapped"], ["PDF"], orc, seed=4, sigma=0.2)


def test_non_periodic_nanoparticle_pdf_with_shape(orc):
    _run_sequence(CASES["ibc_nanoparticle"], ["PDF", "PCF"], orc, seed=5, sigma=0.3, with_shape=True, scale=1.07)


def test_lattice_and_coincident_atoms(orc):
    _run_sequence(CASES["lattice_half"], ["PDF"], orc, seed=6, sigma=0.05)
    _run_sequence(CASES["coincident_empty_class"], ["PDF", "SQ"], orc, seed=7)


def test_five_elements_cfg4_like(orc):
    _run_sequence(CASES["cfg4_small"], ["PDF", "SQ"], orc, seed=8, n_moves=9)


def test_wrap_mode_switch_when_a_move_leaves_the_unit_cell(orc):
    """fast wrap (|frac diff| < 1.5) must hand over to the general wrap when accepted moves drift"""
    case = dict(CASES["ortho_atomic"])
    rng = np.random.default_rng(9)
    store, oracles = _build(case, ["PDF"], rng)
    kw = _hist_kw(case)
    mol, el = case["moleculeIndex"], case["elementIndex"]
    box = case["boxCoords"].copy()
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)
    store.compute_data()
    data_i, data_e = orc.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=el, **kw)
    for shift in (0.3, 0.9, 2.4, -3.1):
        idx = np.array([int(rng.integers(0, box.shape[0]))], dtype=np.int32)
        moved = (box[idx] + F32(shift)).astype(F32)
        bi, be = ep.move_delta(fns, idx, box, kw["basis"], kw["isPBC"], mol, el, kw["numberOfElements"],
                               kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
        tmp = box.copy(); tmp[idx] = moved
        ai, ae = ep.move_delta(fns, idx, tmp, kw["basis"], kw["isPBC"], mol, el, kw["numberOfElements"],
                               kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
        data_i, data_e, box = data_i - bi + ai, data_e - be + ae, tmp
        chi2 = store.propose(idx, moved)
        _check_models(store, oracles, data_i, data_e, chi2, staged=True)
        store.accept()
    gi, ge = store.export_data(0)
    assert np.array_equal(gi, data_i) and np.array_equal(ge, data_e)
    # and a full recomputation on the drifted coordinates uses the general wrap too
    store.compute_data()
    fi, fe = orc.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=el, **kw)
    gi, ge = store.export_data(0)
    assert np.array_equal(gi, fi) and np.array_equal(ge, fe)
    store.close()


def _large_sparse_case(kind):
    """many blocks of the k-d ordered store and a maxDistance far below the cell size (the regime in which the
    full-histogram kernel culls most block pairs); molecules of 5, two elements"""
    rng = np.random.default_rng(404)
    n = 26000
    el = rng.integers(0, 2, n)
    mol = np.arange(n) // 5
    if kind == "ortho":
        return C._case("sparse_ortho", rng.random((n, 3), dtype=F32), np.diag([110.0, 104.0, 98.0]), True, mol, el, 2, 0.0, 6.0, 0.05, 120)
    if kind == "tri_unwrapped":
        box = (rng.random((n, 3), dtype=F32) * F32(2.6) - F32(0.8)).astype(F32)
        return C._case("sparse_tri", box, np.array([[110, 0, 0], [21, 100, 0], [-17, 25, 96]]), True, mol, el, 2, 0.3, 6.3, 0.05, 120)
    box = (rng.random((n, 3), dtype=F32) * F32(100.0) - F32(30.0)).astype(F32)
    return C._case("sparse_ibc", box, np.eye(3), False, mol, el, 2, 0.0, 6.0, 0.05, 120)


@pytest.mark.parametrize("kind", ["ortho", "tri_unwrapped", "non_periodic"])
def test_move_sequences_on_large_sparse_systems(kind, orc):
    """long jumps, seam crossings and molecule moves in a store whose compute_data culls most block pairs: the
    running histograms and chi^2 must stay identical to the reference sequence, through both resolve paths"""
    case = _large_sparse_case(kind)
    rng = np.random.default_rng(9)
    store, oracles = _build(case, ["PDF", "SQ"], rng)
    kw = _hist_kw(case)
    mol, el = case["moleculeIndex"], case["elementIndex"]
    box = case["boxCoords"].copy()
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)
    store.compute_data()
    data_i, data_e = orc.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=el, ncores=orc.max_threads(), **kw)
    span = F32(1.0) if case["isPBC"] else F32(100.0)
    previous = None
    for step in range(36):
        idx = C.group_for(case, rng)
        jump = (0.45 if step % 4 == 0 else 0.004) * span            # every fourth move is a long jump (often across the seam)
        moved = (box[idx] + rng.normal(0, jump, (1, 3)).astype(F32)).astype(F32)
        args = (kw["basis"], kw["isPBC"], mol, el, kw["numberOfElements"], kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
        bi, be = ep.move_delta(fns, idx, box, *args)
        tmp = box.copy(); tmp[idx] = moved
        ai, ae = ep.move_delta(fns, idx, tmp, *args)
        new_i, new_e = data_i - bi + ai, data_e - be + ae
        if step < 18:
            chi2 = store.propose(idx, moved)                         # explicit resolve kernels, state exported every step
            _check_models(store, oracles, new_i, new_e, chi2, staged=True)
            accept = step % 3 != 2
            (store.accept if accept else store.reject)()
            if accept:
                data_i, data_e, box = new_i, new_e, tmp
            gi, ge = store.export_data(0)
            assert np.array_equal(gi, data_i) and np.array_equal(ge, data_e), "running histograms diverged at step %d" % step
        else:
            chi2 = store.step(previous, idx, moved).copy()           # fused path: the resolve rides in the next launch
            for m, (total, exp, dw) in enumerate(oracles):
                assert F32(chi2[m]) == F32(ep.standard_error(exp, total(new_i, new_e), dw)), "step %d model %d" % (step, m)
            previous = step % 3 != 2
            if previous:
                data_i, data_e, box = new_i, new_e, tmp
    (store.accept if previous else store.reject)()
    gi, ge = store.export_data(0)
    assert np.array_equal(gi, data_i) and np.array_equal(ge, data_e)
    assert np.array_equal(store.get_coords(), box)
    store.close()


@pytest.fixture
def persistent_mode():
    global PERSISTENT
    PERSISTENT = True
    yield
    PERSISTENT = False


def test_persistent_kernel_reference_sequences(orc, persistent_mode):
    """the same sequences through the persistent per-move kernel (one cooperative launch serving a run of
    proposals, commands through mapped pinned memory): every check of _run_sequence exports the state, so the
    kernel is stopped and restarted after every move -- the lifecycle path"""
    _run_sequence(CASES["ortho_atomic"], ["PDF", "SQ"], orc, n_moves=9, seed=21)
    _run_sequence(CASES["tri_molecular"], ["PCF", "RSQ"], orc, n_moves=9, seed=22, with_weights=True)
    _run_sequence(CASES["ibc_nanoparticle"], ["PDF"], orc, n_moves=6, seed=23, with_shape=True)


@pytest.mark.parametrize("kind", ["ortho", "tri_unwrapped"])
def test_persistent_kernel_long_run(kind, orc, persistent_mode):
    """a run of proposals served by ONE resident kernel (no state export in between): chi^2 of every step and the
    final state equal the reference sequence; the kernel is started once or twice, not once per move"""
    import time
    case = _large_sparse_case(kind)
    rng = np.random.default_rng(31)
    store, oracles = _build(case, ["PDF", "SQ"], rng)
    kw = _hist_kw(case)
    mol, el = case["moleculeIndex"], case["elementIndex"]
    box = case["boxCoords"].copy()
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)
    store.compute_data()
    data_i, data_e = orc.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=el, ncores=orc.max_threads(), **kw)
    args = (kw["basis"], kw["isPBC"], mol, el, kw["numberOfElements"], kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
    moves = []
    for step in range(40):
        idx = C.group_for(case, rng)
        moves.append((idx, rng.normal(0, 0.45 if step % 5 == 0 else 0.004, (1, 3)).astype(F32)))
    # device first, back to back (the oracle in between would let the kernel's idle watchdog fire every move)
    chis, previous = [], None
    dbox = box.copy()
    for step, (idx, shift) in enumerate(moves):
        moved = (dbox[idx] + shift).astype(F32)
        chis.append(store.step(previous, idx, moved).copy())
        previous = step % 3 != 2
        if previous:
            dbox[idx] = moved
        if step == 25:
            time.sleep(0.05)                                 # longer than the idle watchdog: the kernel leaves and is restarted
    (store.accept if previous else store.reject)()
    started, served = store.persistent_stats()
    assert served == 40 and 1 <= started <= 10, (started, served)      # restarts: the sleep, wrap-mode switches of long jumps
    for step, (idx, shift) in enumerate(moves):
        moved = (box[idx] + shift).astype(F32)
        bi, be = ep.move_delta(fns, idx, box, *args)
        tmp = box.copy(); tmp[idx] = moved
        ai, ae = ep.move_delta(fns, idx, tmp, *args)
        new_i, new_e = data_i - bi + ai, data_e - be + ae
        for m, (total, exp, dw) in enumerate(oracles):
            assert F32(chis[step][m]) == F32(ep.standard_error(exp, total(new_i, new_e), dw)), "step %d model %d" % (step, m)
        if step % 3 != 2:
            data_i, data_e, box = new_i, new_e, tmp
    gi, ge = store.export_data(0)
    assert np.array_equal(gi, data_i) and np.array_equal(ge, data_e)
    assert np.array_equal(store.get_coords(), box)
    store.close()


def test_two_persistent_stores_take_turns(orc, persistent_mode):
    """two stores with resident kernels on one GPU cannot run at once (each kernel fills the device): the one
    that is waiting leaves on its idle watchdog and is restarted later; results stay those of the reference"""
    case = CASES["ortho_atomic"]
    rng = np.random.default_rng(77)
    stores = [_build(case, ["PDF"], rng) for _ in range(2)]
    kw = _hist_kw(case)
    mol, el = case["moleculeIndex"], case["elementIndex"]
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)
    args = (kw["basis"], kw["isPBC"], mol, el, kw["numberOfElements"], kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
    box = case["boxCoords"]
    data_i, data_e = orc.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=el, **kw)
    for st, _ in stores:
        st.compute_data()
    for step in range(6):
        idx = C.group_for(case, rng)
        moved = (box[idx] + rng.normal(0, 0.02, (idx.shape[0], 3)).astype(F32)).astype(F32)
        bi, be = ep.move_delta(fns, idx, box, *args)
        tmp = box.copy(); tmp[idx] = moved
        ai, ae = ep.move_delta(fns, idx, tmp, *args)
        for st, oracles in stores:                              # the same move on both stores, alternating
            chi2 = st.step(None, idx, moved).copy()
            total, exp, dw = oracles[0]
            assert F32(chi2[0]) == F32(ep.standard_error(exp, total(data_i - bi + ai, data_e - be + ae), dw))
            st.reject()
    for st, _ in stores:
        st.close()


@pytest.mark.parametrize("kinds", [("PDF", "SQ"), ("PCF", "RSQ")])
def test_window_function_and_multiframe_prior(kinds, orc):
    """the optional tail of the model totals: total = convolve(prior + weight * total, window, "same")
    (Core/Constraint.py:1160-1177, PairDistributionConstraints.py:890-893).  numpy's convolution fixes no summation
    order, so this branch is held to 1e-6 (the north-star tolerance), not to bit equality."""
    case = CASES["ortho_atomic"]
    rng = np.random.default_rng(3)
    store, oracles = _build(case, list(kinds), rng, scale=0.93)
    tails = []
    for m, (total, exp, dw) in enumerate(oracles):
        n_out = exp.shape[0]
        window = np.hanning(9 + 2 * m).astype(F32); window /= np.sum(window)
        prior = (0.05 * rng.standard_normal(n_out)).astype(F32)
        weight = F32(0.6 + 0.1 * m)
        store.set_multiframe_prior(m, prior, weight)
        store.set_window_function(m, window)
        tails.append((prior, weight, window))
    kw = _hist_kw(case)
    mol, el = case["moleculeIndex"], case["elementIndex"]
    box = case["boxCoords"].copy()
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)
    args = (kw["basis"], kw["isPBC"], mol, el, kw["numberOfElements"], kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
    data_i, data_e = orc.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=el, **kw)

    def check(chi2, intra, inter, staged):
        for m, (total, exp, dw) in enumerate(oracles):
            prior, weight, window = tails[m]
            ref = ep.apply_prior_and_window(total(intra, inter), prior, weight, window)
            got = store.export_total(m, staged=staged)
            scale = float(np.max(np.abs(ref)))
            assert np.max(np.abs(got - ref)) <= 1e-6 * scale, "model %d total" % m
            ref_chi = float(ep.standard_error(exp, ref, dw))
            assert abs(float(chi2[m]) - ref_chi) <= 2e-6 * abs(ref_chi), "model %d chi2" % m

    check(store.compute_data(), data_i, data_e, staged=False)
    for step in range(4):
        idx = C.group_for(case, rng)
        moved = (box[idx] + rng.normal(0, 0.02, (idx.shape[0], 3)).astype(F32)).astype(F32)
        bi, be = ep.move_delta(fns, idx, box, *args)
        tmp = box.copy(); tmp[idx] = moved
        ai, ae = ep.move_delta(fns, idx, tmp, *args)
        chi2 = store.propose(idx, moved)
        check(chi2, data_i - bi + ai, data_e - be + ae, staged=True)
        store.accept()
        data_i, data_e, box = data_i - bi + ai, data_e - be + ae, tmp
    # switching both off restores the bit-exact totals
    for m in range(len(oracles)):
        store.set_multiframe_prior(m, None, 0.0)
        store.set_window_function(m, None)
    chi2 = store.compute_data()
    _check_models(store, oracles, data_i, data_e, chi2, staged=False)
    store.close()


def test_state_machine_errors():
    from fullrmc_b200.store import DeviceStore
    case = CASES["tiny_13"]
    store = DeviceStore(case["boxCoords"], case["basis"], True, case["moleculeIndex"], case["elementIndex"], 3)
    store.add_grid(case["minDistance"], case["maxDistance"], case["bin"], case["histSize"])
    idx = np.array([1], dtype=np.int32)
    with pytest.raises(RuntimeError):
        store.propose(idx, case["boxCoords"][idx])          # compute_data first
    store.compute_data()
    with pytest.raises(RuntimeError):
        store.accept()                                      # nothing staged
    store.propose(idx, case["boxCoords"][idx])
    with pytest.raises(RuntimeError):
        store.propose(idx, case["boxCoords"][idx])          # already staged
    store.reject()
    with pytest.raises(ValueError):
        store.propose(np.array([99], dtype=np.int32), case["boxCoords"][idx])
    store.close()


def test_two_grids_one_pass(orc):
    """NiTi-like setup: a fine PDF grid and a coarse S(Q) grid fed by the same pass over the store"""
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    case = CASES["ortho_atomic"]
    rng = np.random.default_rng(10)
    els, n_per, wdict, volume, rho0 = _system_meta(case)
    store = DeviceStore(case["boxCoords"], case["basis"], True, case["moleculeIndex"], case["elementIndex"], 3)
    grids = []
    for (rmin, b, hs) in ((F32(0.005), F32(0.01), 1200), (F32(0.2856), F32(0.2), 60)):
        edges = (rmin + b * np.arange(hs + 1, dtype=np.float64)).astype(F32)
        centers = ((edges[:-1] + edges[1:]) / F32(2.)).astype(F32)
        grids.append(dict(rmin=edges[0], rmax=edges[-1], bin=b, hs=hs, centers=centers, sv=ep.shell_arrays_from_edges(edges)))
    q = np.linspace(0.6, 12.0, 64).astype(F32)
    exps = [rng.normal(0, 0.4, 1200).astype(F32), rng.normal(0, 0.4, 64).astype(F32)]
    for gi_, (g, kind) in enumerate(zip(grids, ("PDF", "RSQ"))):
        gid = store.add_grid(g["rmin"], g["rmax"], g["bin"], g["hs"])
        store.add_model(gid, ModelSpec(kind, els, n_per, wdict, volume, rho0, g["centers"], g["sv"], exps[gi_],
                                       q_values=q if kind == "RSQ" else None))
    chi2 = store.compute_data()
    box = case["boxCoords"].copy()
    mol, el = case["moleculeIndex"], case["elementIndex"]

    def oracle_chi2(boxc):
        out = []
        for gi_, (g, kind) in enumerate(zip(grids, ("PDF", "RSQ"))):
            hi, he = orc.full_pairs_histograms_coords(boxCoords=boxc, basis=case["basis"], isPBC=True, moleculeIndex=mol,
                                                      elementIndex=el, numberOfElements=3, minDistance=g["rmin"],
                                                      maxDistance=g["rmax"], bin=g["bin"], histSize=g["hs"])
            common = dict(elements=els, n_per_element=n_per, weighting=wdict, volume=volume, rho0=rho0,
                          shell_centers=g["centers"], shell_volumes=g["sv"])
            tot = ep.total_Gr(hi, he, **common) if kind == "PDF" else \
                ep.total_Sq(hi, he, gr2sq=ep.gr2sq_matrix(q, g["centers"]), reduced=True, **common)
            out.append(ep.standard_error(exps[gi_], tot))
        return out
    want = oracle_chi2(box)
    assert F32(chi2[0]) == F32(want[0]) and F32(chi2[1]) == F32(want[1])
    for step in range(4):
        idx = np.array([int(rng.integers(0, box.shape[0]))], dtype=np.int32)
        moved = (box[idx] + rng.normal(0, 0.03, (1, 3)).astype(F32)).astype(F32)
        chi2 = store.propose(idx, moved)
        tmp = box.copy(); tmp[idx] = moved
        want = oracle_chi2(tmp)       # k=1: ordered running state == from-scratch state
        assert F32(chi2[0]) == F32(want[0]) and F32(chi2[1]) == F32(want[1])
        store.accept(); box = tmp
    store.close()


def test_constraint_mirrors_five_method_protocol(orc):
    """DevicePairDistributionConstraint + DeviceReducedStructureFactorConstraint driven exactly like
    Engine.__on_runtime_step_try_move drives the reference constraints (Engine.py:3302-3338):
    before/after for every constraint, Metropolis on the summed chi^2, accept_move/reject_move on all."""
    from fullrmc_b200.constraints import (DeviceBackend, DevicePairDistributionConstraint,
                                          DeviceReducedStructureFactorConstraint)
    case = CASES["ortho_atomic"]
    rng = np.random.default_rng(21)
    els, n_per, wdict, volume, rho0 = _system_meta(case)
    box = case["boxCoords"].copy()
    mol, el = case["moleculeIndex"], case["elementIndex"]
    backend = DeviceBackend(box, case["basis"], True, mol, el, els, n_per, volume, rho0)
    # PDF from an "experimental" r column (bin 0.02), S(Q)-1 on its own coarse r-grid (NiTi-like)
    r = (0.01 + 0.02 * np.arange(650)).astype(F32)
    exp_pdf = np.stack([r, rng.normal(0, 0.3, 650).astype(F32)], axis=1).astype(F32)
    q = np.linspace(0.5, 15.0, 117).astype(F32)
    exp_sq = np.stack([q, rng.normal(0, 0.2, 117).astype(F32)], axis=1).astype(F32)
    pdf = DevicePairDistributionConstraint(backend, exp_pdf, wdict)
    rsf = DeviceReducedStructureFactorConstraint(backend, exp_sq, wdict, rmin=0.3, rmax=14.0, dr=0.2)
    constraints = [pdf, rsf]

    def oracle(boxc):
        out = []
        for c, kind in ((pdf, "PDF"), (rsf, "RSQ")):
            hi, he = orc.full_pairs_histograms_coords(boxCoords=boxc, basis=case["basis"], isPBC=True, moleculeIndex=mol,
                                                      elementIndex=el, numberOfElements=3, minDistance=c.minimumDistance,
                                                      maxDistance=c.maximumDistance, bin=c.bin, histSize=c.histogramSize)
            common = dict(elements=els, n_per_element=n_per, weighting=wdict, volume=volume, rho0=rho0,
                          shell_centers=c.shellCenters, shell_volumes=c.shellVolumes)
            if kind == "PDF":
                tot = ep.total_Gr(hi, he, **common)
            else:
                tot = ep.total_Sq(hi, he, gr2sq=ep.gr2sq_matrix(q, c.shellCenters), reduced=True, **common)
            out.append(ep.standard_error(c.experimentalData, tot))
        return out

    for c in constraints:
        data, err = c.compute_data()
    want = oracle(box)
    assert F32(pdf.standardError) == F32(want[0]) and F32(rsf.standardError) == F32(want[1])
    assert pdf.data["inter"].sum() > 0
    total_old = float(pdf.standardError) + float(rsf.standardError)
    n_acc = 0
    for step in range(8):
        idx = np.array([int(rng.integers(0, box.shape[0]))], dtype=np.int32)
        moved = (box[idx] + rng.normal(0, 0.02, (1, 3)).astype(F32)).astype(F32)
        for c in constraints:
            c.compute_before_move(idx, idx)
            c.compute_after_move(idx, idx, moved)
        tmp = box.copy(); tmp[idx] = moved
        want = oracle(tmp)
        assert F32(pdf.afterMoveStandardError) == F32(want[0])
        assert F32(rsf.afterMoveStandardError) == F32(want[1])
        total_new = float(pdf.afterMoveStandardError) + float(rsf.afterMoveStandardError)
        if total_new <= total_old:
            for c in constraints:
                c.accept_move(idx, idx)
            box, total_old = tmp, total_new
            n_acc += 1
        else:
            for c in constraints:
                c.reject_move(idx, idx)
        assert F32(pdf.standardError) == F32(oracle(box)[0])
    assert pdf.tried == 8 and rsf.tried == 8 and pdf.accepted == n_acc
    assert np.array_equal(backend.store.get_coords(), box)
    backend.close()
