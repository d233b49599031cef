"""Golden trajectories recorded from the UNMODIFIED reference constraint classes
(tests/gen_golden_constraints.py: PairDistributionConstraint, PairCorrelationConstraint,
StructureFactorConstraint, ReducedStructureFactorConstraint on the shipped NiTi / THF / SiOx
example inputs and a synthetic triclinic system), replayed

* on the CPU through the oracle restatement (pins oracle/epilogue.py and the M-F sequence
  against the reference's own class code), and
* on the GPU through the device constraint mirrors (the parity test of the stateful path).

Bar: chi^2 after every move, the committed chi^2, totals and the final data arrays bit-exact.

Edge bins: in the NiTi example a Ni-Ni pair lands within an ulp of the S(Q) grid's maxDistance at
move 9, its fp32 bin index equals histSize and the reference's unchecked write spills into the
next slab ([ni,ti], bin 0).  These replays therefore run with the reference-compatible spill
switched on (oracle.set_emulate_spill / fullrmc_b200.set_edge_spill); the default policy (drop and
count) is exercised in tests/test_gpu_stateless.py.
"""
import os

import numpy as np
import pytest

from oracle import epilogue as ep

F32 = np.float32
NAMES = ["niti", "thf", "siox", "synth", "niti_sf", "synth_sf", "siox_shape"]   # *_sf: scale-factor refit; *_shape: shape function refresh
Z = {"o": 8.0, "si": 14.0, "ti": 22.0, "ni": 28.0, "zr": 40.0, "c": 6.0, "h": 1.0}


class _Golden(dict):
    """the fixture as a dict with the NpzFile's ``files`` attribute"""

    @property
    def files(self):
        return list(self.keys())


def _load(golden_dir, name):
    """A trajectory fixture.  Fixtures of the large synthetic systems (tests/gen_golden_large.py) hold the recipe of
    fullrmc_b200.synthetic instead of the per-atom arrays; those, and the final coordinates, are rebuilt here."""
    z = np.load(os.path.join(golden_dir, "constraints_%s.npz" % name))
    g = _Golden((k, z[k]) for k in z.files)
    if "recipe_name" in g:
        from fullrmc_b200 import synthetic
        system = getattr(synthetic, str(g["recipe_name"]))(int(g["recipe_n"]), int(g["recipe_seed"]))
        assert np.array_equal(system.basis, g["basis"])
        g["boxCoords"], g["moleculeIndex"], g["elementIndex"] = system.boxCoords, system.moleculeIndex, system.elementIndex
        final = system.boxCoords.copy()
        for s in range(g["steps/idx"].shape[0]):
            if bool(g["steps/accepted"][s]):
                k = int(g["steps/k"][s])
                final[g["steps/idx"][s, :k]] = g["steps/moved"][s, :k]          # Engine.py:3337-3338
        g["final_boxCoords"] = final
    return g


def _constraint_desc(g, ci):
    d = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("c%d/" % ci)}
    d["kind"] = str(d["kind"])
    d["weighting"] = {str(p): F32(w) for p, w in zip(d["pairs"], d["pair_w"])}
    d["dataWeights"] = None if d["dataWeights"].shape[0] == 0 else d["dataWeights"]
    d["shapeArray"] = None if d["shapeArray"].shape[0] == 0 else d["shapeArray"]
    sp = d.get("shapeFuncParams")
    d["shapeParams"] = None if sp is None else dict(rmin=sp[0], rmax=None if np.isnan(sp[1]) else sp[1], dr=sp[2], qmin=sp[3],
                                                    qmax=sp[4], dq=sp[5], updateFreq=int(d["shapeUpdateFreq"]))
    adj = d.get("adjustScaleFactor")
    d["adjust"] = (0, 0.0, 0.0) if adj is None else (int(adj[0]), F32(adj[1]), F32(adj[2]))
    return d


def _system(g):
    elements = [str(e) for e in g["elements"]]
    counts = np.bincount(g["elementIndex"], minlength=len(elements))
    n_per = {elements[i]: int(counts[i]) for i in range(len(elements))}
    return elements, n_per


def _oracle_total(d, intra, inter, elements, n_per, volume, rho0, sf=None, accepted=None):
    """(total, scale factor used): the committed factor sf, or a refit when the constraint adjusts its scale
    factor and accepted % frequency == 0 (get_adjusted_scale_factor, Core/Constraint.py:1397-1423)"""
    common = dict(elements=elements, n_per_element=n_per, weighting=d["weighting"], volume=volume, rho0=rho0,
                  shell_centers=d["shellCenters"], shell_volumes=d["shellVolumes"])
    sf = float(d["scaleFactor"]) if sf is None else float(sf)
    freq, lo, hi = d["adjust"]
    refit = (d["experimental"], d["dataWeights"], lo, hi) if (freq and accepted is not None and accepted % freq == 0) else None
    if d["kind"] == "PDF":
        out = ep.total_Gr(intra, inter, shape_array=d["shapeArray"], scale_factor=sf, refit=refit, **common)
    elif d["kind"] == "PCF":
        out = ep.total_gr(intra, inter, shape_array=d["shapeArray"], scale_factor=sf, refit=refit, **common)
    else:
        out = ep.total_Sq(intra, inter, gr2sq=ep.gr2sq_matrix(d["qValues"], d["shellCenters"]), scale_factor=sf,
                          reduced=(d["kind"] == "RSQ"), refit=refit, **common)
    return out if refit is not None else (out, F32(sf))


@pytest.fixture
def spill_oracle(orc):
    orc.set_emulate_spill(True)
    yield orc
    orc.set_emulate_spill(False)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_restatement_reproduces_reference_classes(name, golden_dir, spill_oracle):
    orc = spill_oracle
    g = _load(golden_dir, name)
    elements, n_per = _system(g)
    volume, rho0 = F32(g["volume"]), F32(g["numberDensity"])
    box = g["boxCoords"].copy()
    basis, pbc, mol, el = g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["elementIndex"]
    nc = int(g["n_constraints"])
    descs = [_constraint_desc(g, ci) for ci in range(nc)]
    fns = (orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords)
    data = []
    for ci, d in enumerate(descs):
        hi, he = orc.full_pairs_histograms_coords(boxCoords=box, basis=basis, isPBC=pbc, moleculeIndex=mol, elementIndex=el,
                                                  numberOfElements=len(elements), minDistance=d["minDistance"],
                                                  maxDistance=d["maxDistance"], bin=d["bin"], histSize=int(d["histSize"]),
                                                  ncores=orc.max_threads())
        assert np.array_equal(hi, d["start_intra"]) and np.array_equal(he, d["start_inter"])
        data.append([hi, he])
        if d["shapeParams"] is not None:
            continue                                       # checked below, once the first shape array exists
        tot, _ = _oracle_total(d, hi, he, elements, n_per, volume, rho0, accepted=0)
        assert np.array_equal(tot, d["start_total"]), "constraint %d total differs from the reference class" % ci
        chi = ep.standard_error(d["experimental"], tot, d["dataWeights"])
        assert F32(chi) == F32(g["start_stdErr"][ci])
    steps = g["steps/idx"].shape[0]
    sfs = [F32(d["scaleFactor"]) for d in descs]          # committed scale factors
    accepted = 0                                           # engine.accepted
    # shape function refreshed from the running configuration (Collection.py:20-125): the numpy restatement
    # oracle/shape.py is pinned here (bit for bit); the product's device kernel is held to it in the GPU test
    from oracle import shape as fshape
    shape_w = ep.normalized_weighting(n_per, {e: Z[e] for e in elements})
    last_shape = {}

    def rebuild_shape(ci, d, coords, k):
        p = d["shapeParams"]
        rmax = p["rmax"] if p["rmax"] is not None else fshape.auto_rmax(pbc, basis, coords)
        arr = fshape.get_Gr_shape_function(d["shellCenters"], coords, basis, pbc, mol, el, elements, n_per, volume, shape_w,
                                           qmin=p["qmin"], qmax=p["qmax"], dq=p["dq"], rmin=p["rmin"], rmax=rmax, dr=p["dr"],
                                           full_histogram=lambda **kw: orc.full_pairs_histograms_coords(ncores=orc.max_threads(), **kw))
        assert np.array_equal(arr, d["shape_arrays"][k]), "constraint %d shape array %d differs from the reference's" % (ci, k)
        d["shapeArray"] = arr
        last_shape[ci] = [accepted, k + 1]

    for ci, d in enumerate(descs):
        if d["shapeParams"] is not None:
            rebuild_shape(ci, d, box, 0)
            tot, _ = _oracle_total(d, data[ci][0], data[ci][1], elements, n_per, volume, rho0, accepted=0)
            assert F32(ep.standard_error(d["experimental"], tot, d["dataWeights"])) == F32(g["start_stdErr"][ci])
    for s in range(steps):
        k = int(g["steps/k"][s])
        idx = g["steps/idx"][s, :k].astype(np.int32)
        moved = g["steps/moved"][s, :k]
        tmp = box.copy(); tmp[idx] = moved
        for ci, d in enumerate(descs):                      # _runtime_on_step (PairDistributionConstraints.py:362-374)
            if d["shapeParams"] is not None and last_shape[ci][0] != accepted and accepted % d["shapeParams"]["updateFreq"] == 0:
                assert int(d["shape_steps"][last_shape[ci][1]]) == s
                rebuild_shape(ci, d, box, last_shape[ci][1])
        staged, used = [], []
        for ci, d in enumerate(descs):
            args = (basis, pbc, mol, el, len(elements), d["minDistance"], d["maxDistance"], d["bin"], int(d["histSize"]))
            bi, be = ep.move_delta(fns, idx, box, *args)
            ai, ae = ep.move_delta(fns, idx, tmp, *args)
            ni, ne = data[ci][0] - bi + ai, data[ci][1] - be + ae
            tot, sf_used = _oracle_total(d, ni, ne, elements, n_per, volume, rho0, sf=sfs[ci], accepted=accepted)
            chi = ep.standard_error(d["experimental"], tot, d["dataWeights"])
            assert F32(chi) == F32(g["steps/chi2_after"][s, ci]), "step %d constraint %d" % (s, ci)
            if "steps/scale_used" in g.files:
                assert F32(sf_used) == F32(g["steps/scale_used"][s, ci]), "step %d constraint %d scale factor" % (s, ci)
            staged.append([ni, ne]); used.append(F32(sf_used))
        if bool(g["steps/accepted"][s]):
            data, box, sfs = staged, tmp, used             # accept_move commits the factor the evaluation used
            accepted += 1
    for ci, d in enumerate(descs):
        assert np.array_equal(data[ci][0], d["final_intra"]) and np.array_equal(data[ci][1], d["final_inter"])
        tot, _ = _oracle_total(d, data[ci][0], data[ci][1], elements, n_per, volume, rho0, sf=sfs[ci], accepted=accepted)
        assert np.array_equal(tot, d["final_total"])
        if "c%d/final_scaleFactor" % ci in g.files:
            assert F32(sfs[ci]) == F32(g["c%d/final_scaleFactor" % ci])
    assert np.array_equal(box, g["final_boxCoords"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_constraints_reproduce_reference_classes(name, golden_dir):
    import fullrmc_b200
    from fullrmc_b200.constraints import DeviceBackend, make_device_constraint
    previous = fullrmc_b200.set_edge_spill(True)
    try:
        _replay_on_device(name, golden_dir, DeviceBackend, make_device_constraint)
    finally:
        fullrmc_b200.set_edge_spill(previous)


def _close(a, b, exact):
    """bit equality, or -- where a device-computed shape function enters (double-precision sums on the device against
    the reference's float32 numpy sums) -- the north-star tolerance: 1e-6 relative, norm-wise for arrays"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if exact:
        return bool(np.array_equal(a, b))
    return bool(np.linalg.norm(a - b) <= 1e-6 * max(np.linalg.norm(b), 1e-30))


def _replay_on_device(name, golden_dir, DeviceBackend, make_device_constraint):
    from fullrmc_b200 import model as fm
    g = _load(golden_dir, name)
    exact = not any(k.endswith("/shapeFuncParams") for k in g.files)
    elements, n_per = _system(g)
    shape_w = fm.faber_ziman_weights(n_per, {e: Z[e] for e in elements})
    backend = DeviceBackend(g["boxCoords"], g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["elementIndex"], elements,
                            n_per, g["volume"], g["numberDensity"])
    nc = int(g["n_constraints"])
    cons = []
    for ci in range(nc):
        d = _constraint_desc(g, ci)
        cons.append((d, make_device_constraint(backend, d["kind"], d["experimental"], d["minDistance"], d["maxDistance"], d["bin"],
                                               int(d["histSize"]), d["shellCenters"], d["shellVolumes"], d["weighting"],
                                               dataWeights=d["dataWeights"], shapeArray=d["shapeArray"],
                                               scaleFactor=float(d["scaleFactor"]),
                                               qValues=d.get("qValues") if d["kind"] in ("SQ", "RSQ") else None,
                                               adjustScaleFactor=d["adjust"], shapeFuncParams=d["shapeParams"],
                                               shapeWeighting=shape_w)))
    n_shapes = {}
    for ci, (d, c) in enumerate(cons):
        data, err = c.compute_data()
        assert np.array_equal(data["intra"], d["start_intra"]) and np.array_equal(data["inter"], d["start_inter"])
        if d["shapeParams"] is not None:                    # Engine.run: _runtime_initialize builds the first shape array
            c.runtime_initialize()
            assert _close(c._shapeArray, d["shape_arrays"][0], exact)
            n_shapes[ci] = 1
            err = c.standardError
        assert _close(c.get_constraint_total(), d["start_total"], exact)
        assert _close(err, g["start_stdErr"][ci], exact)
    steps = g["steps/idx"].shape[0]
    for s in range(steps):
        k = int(g["steps/k"][s])
        idx = g["steps/idx"][s, :k].astype(np.int32)
        moved = np.ascontiguousarray(g["steps/moved"][s, :k])
        for ci, (d, c) in enumerate(cons):                  # Engine.run: _runtime_on_step before every move
            if d["shapeParams"] is not None and c.runtime_on_step():
                assert int(d["shape_steps"][n_shapes[ci]]) == s
                assert _close(c._shapeArray, d["shape_arrays"][n_shapes[ci]], exact), "shape array %d" % n_shapes[ci]
                n_shapes[ci] += 1
        for d, c in cons:
            c.compute_before_move(idx, idx)
            c.compute_after_move(idx, idx, moved)
        for ci, (d, c) in enumerate(cons):
            assert _close(c.afterMoveStandardError, g["steps/chi2_after"][s, ci], exact), "step %d constraint %d" % (s, ci)
            if "steps/scale_used" in g.files:
                assert F32(c.fittedScaleFactor) == F32(g["steps/scale_used"][s, ci]), "step %d constraint %d scale factor" % (s, ci)
        for d, c in cons:
            (c.accept_move if bool(g["steps/accepted"][s]) else c.reject_move)(idx, idx)
    for ci, n in n_shapes.items():
        assert n == len(cons[ci][0]["shape_steps"])          # every refresh of the reference happened here too
    refits = any(d["adjust"][0] for d, _ in cons)
    for ci, (d, c) in enumerate(cons):
        data = c.data
        assert np.array_equal(data["intra"], d["final_intra"]) and np.array_equal(data["inter"], d["final_inter"])
        assert _close(c.standardError, d["final_stdErr"], exact)
        if "c%d/final_scaleFactor" % ci in g.files:
            assert F32(c.scaleFactor) == F32(g["c%d/final_scaleFactor" % ci])
        if refits:
            # the recorded final total is a fresh evaluation at the final accepted count (it may refit): do the same
            c.compute_data(update=False)
        assert _close(c.get_constraint_total(), d["final_total"], exact)
    assert np.array_equal(backend.store.get_coords(), g["final_boxCoords"])
    backend.close()


def test_structure_factor_constraint_builds_the_reference_grid(golden_dir):
    """DeviceStructureFactorConstraint(rmin, rmax, dr) derives the r-grid the reference class derives from the same
    arguments (round-1 advice: arange(rmin, rmax + dr, dr) when rmax is given, half the shortest basis vector when not).
    The fixtures hold what the unmodified classes computed: NiTi's reduced S(Q) is built with all three left to their
    defaults, the large synthetic one with (0, 19.98, 0.02).  No device is needed: the grid is host arithmetic."""
    from fullrmc_b200 import constraints as fc

    class Backend(object):                      # registration only records the grid
        def __init__(self, basis):
            self.basisVectors = np.asarray(basis, F32); self.elements = []; self.numberOfAtomsPerElement = {}
            self.volume = F32(1); self.numberDensity = F32(1)
        def _register(self, c, grid_key, spec):
            self.grid_key = grid_key

    import unittest.mock as mock
    for name, ci, args in (("niti", 1, {}), ("cfg4", 1, dict(rmin=0.0, rmax=19.98, dr=0.02)), ("synth", 1, {})):
        path = os.path.join(golden_dir, "constraints_%s.npz" % name)
        if not os.path.exists(path):
            continue
        z = np.load(path)
        d = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("c%d/" % ci)}
        if name == "synth":
            args = dict(rmin=None, rmax=None, dr=None)
        exp = np.stack([d["qValues"], d["experimental"]], 1).astype(F32)
        b = Backend(z["basis"])
        with mock.patch.object(fc, "ModelSpec", lambda *a, **k: None):
            c = fc.DeviceStructureFactorConstraint(b, exp, {}, **args)
        assert c.histogramSize == int(d["histSize"]), name
        assert F32(c.minimumDistance) == F32(d["minDistance"]) and F32(c.maximumDistance) == F32(d["maxDistance"]) and F32(c.bin) == F32(d["bin"]), name
        assert np.array_equal(c.shellCenters, d["shellCenters"]) and np.array_equal(c.shellVolumes, d["shellVolumes"]), name
