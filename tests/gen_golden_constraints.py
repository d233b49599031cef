"""Generates tests/golden/constraints_<case>.npz by running the UNMODIFIED reference constraint
classes (PairDistributionConstraint, PairCorrelationConstraint, StructureFactorConstraint,
ReducedStructureFactorConstraint from /root/reference) with the reference's own compiled kernels,
on the shipped example inputs (BASELINE.json configs 1-3) and a synthetic triclinic system.

Run in the build container:   python tests/gen_golden_constraints.py

Each fixture holds the engine arrays, everything the constraint derived from its experimental data
(limits, bin, histogram size, shell centres/volumes, weighting scheme, data weights), and a Metropolis
trajectory driven exactly like Engine.__on_runtime_step_try_move (Engine.py:3302-3338): per step the
moved group, the moved box coordinates, every constraint's afterMoveStandardError, the decision, plus
the final data["intra"/"inter"] arrays and totals.  tests/test_golden_constraints.py replays them
through the oracle restatement (CPU) and through the CUDA path (GPU).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

EX = H.examples_dir()


def read_pdb(path):
    """fixed-column ATOM reader (what Engine.set_pdb takes from pdbparser): coordinates, element,
    molecule key (residue name, sequence number, segment id) and the REMARK Boundary Conditions line"""
    xyz, els, keys, basis = [], [], [], None
    for line in open(path):
        if line.startswith("REMARK") and "Boundary Conditions:" in line:
            v = [float(x) for x in line.split("Boundary Conditions:")[1].split()]
            if len(v) == 9:
                basis = np.array(v, dtype=np.float64).reshape(3, 3)
        if line.startswith(("ATOM", "HETATM")):
            xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            els.append(line[76:78].strip().lower())
            keys.append((line[17:21].strip(), line[22:26].strip(), line[72:76].strip()))
    mol, last, cur = [], None, -1
    for k in keys:
        if k != last:
            cur += 1; last = k
        mol.append(cur)
    return np.array(xyz, dtype=np.float32), els, np.array(mol, dtype=np.int32), basis


def engine_arrays(xyz, els, mol, basis):
    from fullrmc.Core.boundary_conditions_collection import transform_coordinates
    elements = sorted(set(els))
    el = np.array([elements.index(e) for e in els], dtype=np.int32)
    if basis is None:
        return xyz.copy(), np.eye(3, dtype=np.float32), False, mol, el, elements
    basis32 = basis.astype(np.float32)
    rbasis = np.linalg.inv(basis).astype(np.float32)
    box = transform_coordinates(transMatrix=rbasis, coords=xyz)       # Engine.py:2265
    return np.ascontiguousarray(box, dtype=np.float32), basis32, True, mol, el, elements


def describe(c, kind):
    """what the reference constraint derived from its inputs"""
    pre = {"PDF": "_PairDistributionConstraint__", "PCF": "_PairDistributionConstraint__",
           "SQ": "_StructureFactorConstraint__", "RSQ": "_StructureFactorConstraint__"}[kind]
    g = lambda name: getattr(c, pre + name)
    d = dict(kind=kind, minDistance=np.float32(g("minimumDistance")), maxDistance=np.float32(g("maximumDistance")),
             bin=np.float32(g("bin")), histSize=np.int32(g("histogramSize")), shellCenters=np.asarray(g("shellCenters"), np.float32),
             shellVolumes=np.asarray(g("shellVolumes"), np.float32), scaleFactor=np.float32(c.scaleFactor))
    pairs = g("elementsPairs")
    ws = g("weightingScheme")
    d["pairs"] = np.array(["%s-%s" % p for p in pairs])
    d["pair_w"] = np.array([ws.get("%s-%s" % p, ws.get("%s-%s" % (p[1], p[0]))) for p in pairs], dtype=np.float32)
    if kind in ("PDF", "PCF"):
        d["experimental"] = np.asarray(c.experimentalPDF, np.float32)
    else:
        d["experimental"] = np.asarray(g("experimentalSF"), np.float32)
        d["qValues"] = np.asarray(g("experimentalQValues"), np.float32)
    d["dataWeights"] = np.zeros(0, np.float32) if c._usedDataWeights is None else np.asarray(c._usedDataWeights, np.float32)
    d["shapeArray"] = np.zeros(0, np.float32) if getattr(c, "_shapeArray", None) is None else np.asarray(c._shapeArray, np.float32)
    d["adjustScaleFactor"] = np.array([c.adjustScaleFactorFrequency, c.adjustScaleFactorMinimum, c.adjustScaleFactorMaximum], np.float64)
    return d


def fit_total(c, kind, E):
    """the array chi^2 is computed from: the constraint's PRIVATE fitting path (__get_total_Gr /
    __get_total_gr / __get_total_Sq), not the plotting path of get_constraint_value (different op order)"""
    fn = {"PDF": "_PairDistributionConstraint__get_total_Gr", "PCF": "_PairCorrelationConstraint__get_total_gr",
          "SQ": "_StructureFactorConstraint__get_total_Sq", "RSQ": "_StructureFactorConstraint__get_total_Sq"}[kind]
    return np.asarray(getattr(c, fn)(c.data, rho0=E.numberDensity), np.float32)


# command line: [--dropin] [--out DIR] [case names: regenerate just those]
#   --dropin   the same unmodified reference classes, but with fullrmc.Core.<extension> replaced by the CUDA drop-in
#              modules fullrmc_b200.Core.<name> (tests/ref_harness.load_reference(dropin=True)); needs a GPU.
#              tests/test_dropin.py runs this into a scratch directory and requires the files it writes to equal the
#              committed fixtures array for array: the literal drop-in demonstration.
_ARGS = sys.argv[1:]
DROPIN = "--dropin" in _ARGS
OUT_DIR = _ARGS[_ARGS.index("--out") + 1] if "--out" in _ARGS else None
ONLY = set(a for i, a in enumerate(_ARGS) if not a.startswith("--") and (i == 0 or _ARGS[i - 1] != "--out"))


def run_case(name, fullrmc, arrays, make_constraints, groups, n_steps, seed, sigma, out_dir, recipe=None):
    """recipe: (generator name, n, seed) of fullrmc_b200.synthetic for systems too large to store; the fixture then
    holds the recipe instead of the per-atom arrays and the tests regenerate them (tests/test_golden_large.py)"""
    if ONLY and name not in ONLY:
        return
    box, basis, isPBC, mol, el, elements = arrays
    E = H.fake_engine(fullrmc, box, basis, isPBC, mol, el, elements)
    constraints = make_constraints(E)
    out = dict(basis=basis, isPBC=np.bool_(isPBC), elements=np.array(elements), volume=np.float32(E.volume),
               numberDensity=np.float32(E.numberDensity), n_constraints=np.int32(len(constraints)))
    if recipe is None:
        out.update(boxCoords=box.copy(), moleculeIndex=mol, elementIndex=el)
    else:
        out.update(recipe_name=np.array(recipe[0]), recipe_n=np.int64(recipe[1]), recipe_seed=np.int64(recipe[2]))
    for ci, (c, kind) in enumerate(constraints):
        H.attach(E, c)
        for k, v in describe(c, kind).items():
            out["c%d/%s" % (ci, k)] = v
    rng = np.random.default_rng(seed)
    start = []
    shaped = [ci for ci, (c, kind) in enumerate(constraints) if getattr(c, "_shapeFuncParams", None) is not None]
    shape_log = {ci: [] for ci in shaped}          # (step at which the array was (re)built, array)
    if shaped:
        object.__setattr__(E, "_Engine__totalStandardError", 0.0)
        object.__setattr__(E, "update_total_standard_error", lambda: None)
    for ci, (c, kind) in enumerate(constraints):
        data, err = c.compute_data()
        start.append(np.float32(c.standardError))
        out["c%d/start_intra" % ci], out["c%d/start_inter" % ci] = data["intra"].copy(), data["inter"].copy()
        out["c%d/start_total" % ci] = fit_total(c, kind, E)
    for ci in shaped:                              # Engine.run: _runtime_initialize builds the first shape array
        c = constraints[ci][0]
        c._runtime_initialize()
        shape_log[ci].append((-1, np.asarray(c._shapeArray, np.float32).copy()))
        start[ci] = np.float32(c.standardError)
        out["c%d/start_total" % ci] = fit_total(c, constraints[ci][1], E)
        out["c%d/shapeFuncParams" % ci] = np.array([c._shapeFuncParams[k] if c._shapeFuncParams[k] is not None else np.nan
                                                    for k in ("rmin", "rmax", "dr", "qmin", "qmax", "dq")], np.float64)
        out["c%d/shapeUpdateFreq" % ci] = np.int32(c._shapeUpdateFreq)
    out["start_stdErr"] = np.array(start, np.float32)
    rbasis = np.linalg.inv(basis.astype(np.float64)) if isPBC else np.eye(3)
    idx_log, moved_log, chi_log, acc_log, k_log, sf_log = [], [], [], [], [], []
    total_old = sum(float(c.standardError) for c, _ in constraints)
    for step in range(n_steps):
        for ci in shaped:                          # Engine.run calls _runtime_on_step before every move
            c = constraints[ci][0]
            before = c._lastShapeUpdate
            c._runtime_on_step()
            if c._lastShapeUpdate != before:
                shape_log[ci].append((step, np.asarray(c._shapeArray, np.float32).copy()))
                total_old = sum(float(cc.standardError) for cc, _ in constraints)
        idx = np.asarray(groups[int(rng.integers(0, len(groups)))], dtype=np.int32)
        shift = (rng.normal(0.0, sigma, (1, 3)) @ rbasis).astype(np.float32)          # rigid translation of the group
        moved = (E.boxCoordinates[idx] + shift).astype(np.float32)
        for c, _ in constraints:
            c.compute_before_move(realIndexes=idx, relativeIndexes=idx)
            c.compute_after_move(realIndexes=idx, relativeIndexes=idx, movedBoxCoordinates=moved)
        chis = [np.float32(c.afterMoveStandardError) for c, _ in constraints]
        sf_log.append([np.float32(c._fittedScaleFactor) for c, _ in constraints])     # the scale factor this evaluation used
        total_new = sum(float(x) for x in chis)
        accept = total_new <= total_old or step % 5 == 4            # also exercise uphill accepts
        for c, _ in constraints:
            (c.accept_move if accept else c.reject_move)(realIndexes=idx, relativeIndexes=idx)
        if accept:
            E.boxCoordinates[idx] = moved                            # Engine.py:3337-3338
            total_old = total_new
            object.__setattr__(E, "_Engine__accepted", E.accepted + 1)
        idx_log.append(np.pad(idx, (0, 64 - idx.shape[0]), constant_values=-1)); k_log.append(idx.shape[0])
        moved_log.append(np.pad(moved, ((0, 64 - idx.shape[0]), (0, 0)))); chi_log.append(chis); acc_log.append(accept)
    out["steps/idx"] = np.array(idx_log, np.int32)
    out["steps/k"] = np.array(k_log, np.int32)
    out["steps/moved"] = np.array(moved_log, np.float32)
    out["steps/chi2_after"] = np.array(chi_log, np.float32)
    out["steps/accepted"] = np.array(acc_log, np.bool_)
    out["steps/scale_used"] = np.array(sf_log, np.float32)
    for ci, (c, kind) in enumerate(constraints):
        out["c%d/final_intra" % ci], out["c%d/final_inter" % ci] = c.data["intra"].copy(), c.data["inter"].copy()
        out["c%d/final_stdErr" % ci] = np.float32(c.standardError)
        out["c%d/final_scaleFactor" % ci] = np.float32(c.scaleFactor)
        out["c%d/final_total" % ci] = fit_total(c, kind, E)
    for ci in shaped:
        out["c%d/shape_steps" % ci] = np.array([t for t, _ in shape_log[ci]], np.int32)
        out["c%d/shape_arrays" % ci] = np.array([a for _, a in shape_log[ci]], np.float32)
    if recipe is None:
        out["final_boxCoords"] = np.asarray(E.boxCoordinates, np.float32).copy()
    path = os.path.join(out_dir, "constraints_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-10s %5d atoms, %d constraints, %d steps (%d accepted), start chi2 %s -> %s   [%d KiB]" % (
        name, box.shape[0], len(constraints), n_steps, int(np.sum(acc_log)), start,
        [float(c.standardError) for c, _ in constraints], os.path.getsize(path) // 1024))


def main():
    fullrmc = H.load_reference(dropin=DROPIN)
    assert fullrmc is not None, "needs /root/reference (or the package staged by oracle/build_ref.py)"
    if DROPIN:
        import fullrmc_b200
        fullrmc_b200.set_edge_spill(True)        # the reference's unchecked write at bin == histSize
    from fullrmc.Globals import FLOAT_TYPE
    from fullrmc.Core.Collection import rebin, convert_Gr_to_gr
    from fullrmc.Constraints.PairDistributionConstraints import PairDistributionConstraint
    from fullrmc.Constraints.PairCorrelationConstraints import PairCorrelationConstraint
    from fullrmc.Constraints.StructureFactorConstraints import StructureFactorConstraint, ReducedStructureFactorConstraint
    out_dir = OUT_DIR or os.path.join(ROOT, "tests", "golden")

    # ---- config 1: Examples/atomicNiTi as shipped (run.py:41-49): PDF + reduced S(Q), atomic groups
    d = os.path.join(EX, "atomicNiTi")
    arrays = engine_arrays(*read_pdb(os.path.join(d, "system.pdb")))
    def niti(E):
        pdf = PairDistributionConstraint(experimentalData=os.path.join(d, "experimental.gr"), weighting="atomicNumber")
        Sq = np.transpose(rebin(np.loadtxt(os.path.join(d, "experimental.fq")), bin=0.05)).astype(FLOAT_TYPE)
        rsf = ReducedStructureFactorConstraint(experimentalData=Sq, weighting="atomicNumber")
        return [(pdf, "PDF"), (rsf, "RSQ")]
    n = arrays[0].shape[0]
    run_case("niti", fullrmc, arrays, niti, [[i] for i in range(n)], 160, 1, 0.15, out_dir)

    # ---- config 2: Examples/molecularTHF as shipped (run.py:52-57): g(r) with data weights, molecule moves
    d = os.path.join(EX, "molecularTHF")
    arrays = engine_arrays(*read_pdb(os.path.join(d, "thf.pdb")))
    def thf(E):
        _, _, _, gr = convert_Gr_to_gr(np.loadtxt(os.path.join(d, "thf_pdf.exp")), minIndex=[4, 5, 6])
        dw = np.ones(gr.shape[0]); dw[:np.nonzero(gr[:, 1] > 0)[0][0]] = 0
        return [(PairCorrelationConstraint(experimentalData=gr.astype(FLOAT_TYPE), weighting="atomicNumber", dataWeights=dw), "PCF")]
    mol = arrays[3]
    groups = [np.flatnonzero(mol == m).tolist() for m in range(int(mol.max()) + 1)]
    run_case("thf", fullrmc, arrays, thf, groups, 72, 2, 0.2, out_dir)

    # ---- config 3: Examples/SiOxNanosphere (run.py:41): non-periodic PDF (shape function left to a later round)
    d = os.path.join(EX, "SiOxNanosphere")
    arrays = engine_arrays(*read_pdb(os.path.join(d, "SiOx.pdb")))
    def siox(E):
        object.__setattr__(E, "_Engine__numberDensity", FLOAT_TYPE(0.0125))          # run.py:61 set_number_density
        object.__setattr__(E, "_Engine__volume", FLOAT_TYPE(E.numberOfAtoms / 0.0125))
        return [(PairDistributionConstraint(experimentalData=os.path.join(d, "SiOx.gr"), weighting="atomicNumber"), "PDF")]
    n = arrays[0].shape[0]
    run_case("siox", fullrmc, arrays, siox, [[i] for i in range(n)], 160, 3, 0.2, out_dir)

    # ---- synthetic triclinic, 4 elements: G(r) + full S(Q) with a scale factor and data weights
    rng = np.random.default_rng(44)
    n = 2400
    box = (rng.random((n, 3), dtype=np.float32) * np.float32(1.6) - np.float32(0.3)).astype(np.float32)   # partly unwrapped
    basis = np.array([[33, 0, 0], [5, 32, 0], [-3.5, 6.5, 31]], dtype=np.float32)
    el = rng.integers(0, 4, n).astype(np.int32)
    mol = (np.arange(n) // 3).astype(np.int32)
    arrays = (box, basis, True, mol, el, ["o", "si", "ti", "zr"])
    def synth(E):
        r = (0.05 + 0.05 * np.arange(300)).astype(np.float32)
        pdf = PairDistributionConstraint(experimentalData=np.stack([r, rng.normal(0, 0.2, 300).astype(np.float32)], 1).astype(np.float32),
                                         weighting="atomicNumber", scaleFactor=0.95, dataWeights=rng.random(300))
        q = np.linspace(0.6, 14.0, 150).astype(np.float32)
        sf = StructureFactorConstraint(experimentalData=np.stack([q, 1 + rng.normal(0, 0.1, 150).astype(np.float32)], 1).astype(np.float32),
                                       weighting="atomicNumber", scaleFactor=1.05)
        return [(pdf, "PDF"), (sf, "SQ")]
    groups = [[3 * m, 3 * m + 1, 3 * m + 2] for m in range(n // 3)]
    run_case("synth", fullrmc, arrays, synth, groups, 120, 4, 0.25, out_dir)

    # ---- scale-factor refit (Core/Constraint.py:1363-1423).  NiTi as shipped switches it on for both constraints
    #      (Examples/atomicNiTi/run.py:102-103); a short frequency makes 40 steps cross several refit windows.
    d = os.path.join(EX, "atomicNiTi")
    arrays_niti = engine_arrays(*read_pdb(os.path.join(d, "system.pdb")))
    def niti_sf(E):
        cons = niti(E)
        for c, _ in cons:
            c.set_adjust_scale_factor((4, 0.8, 1.2))
        return cons
    nn = arrays_niti[0].shape[0]
    run_case("niti_sf", fullrmc, arrays_niti, niti_sf, [[i] for i in range(nn)], 160, 11, 0.15, out_dir)

    # ---- config 3 with its shape function (Examples/SiOxNanosphere/run.py:41-47; Constraints/Collection.py:20-125),
    #      refreshed every 4 accepted moves instead of every 1000
    d = os.path.join(EX, "SiOxNanosphere")
    arrays_siox = engine_arrays(*read_pdb(os.path.join(d, "SiOx.pdb")))
    def siox_shape(E):
        object.__setattr__(E, "_Engine__numberDensity", FLOAT_TYPE(0.0125))
        object.__setattr__(E, "_Engine__volume", FLOAT_TYPE(E.numberOfAtoms / 0.0125))
        pdf = PairDistributionConstraint(experimentalData=os.path.join(d, "SiOx.gr"), weighting="atomicNumber")
        pdf.set_shape_function_parameters({'rmin': 0., 'rmax': None, 'dr': 0.5, 'qmin': 0.0001, 'qmax': 0.6, 'dq': 0.005,
                                           'updateFreq': 4})
        return [(pdf, "PDF")]
    ns = arrays_siox[0].shape[0]
    run_case("siox_shape", fullrmc, arrays_siox, siox_shape, [[i] for i in range(ns)], 96, 13, 0.2, out_dir)

    # synthetic: g(r) with data weights + full S(Q), refit every 3 accepted moves with a tight clip range
    rng2 = np.random.default_rng(45)
    def synth_sf(E):
        r = (0.05 + 0.05 * np.arange(300)).astype(np.float32)
        pcf = PairCorrelationConstraint(experimentalData=np.stack([r, 1 + rng2.normal(0, 0.2, 300).astype(np.float32)], 1).astype(np.float32),
                                        weighting="atomicNumber", scaleFactor=0.97, dataWeights=rng2.random(300),
                                        adjustScaleFactor=(3, 0.9, 1.02))
        q = np.linspace(0.6, 14.0, 150).astype(np.float32)
        sf = StructureFactorConstraint(experimentalData=np.stack([q, 1 + rng2.normal(0, 0.1, 150).astype(np.float32)], 1).astype(np.float32),
                                       weighting="atomicNumber", scaleFactor=1.05, adjustScaleFactor=(3, 0.7, 1.3))
        return [(pcf, "PCF"), (sf, "SQ")]
    run_case("synth_sf", fullrmc, arrays, synth_sf, groups, 120, 5, 0.25, out_dir)


if __name__ == "__main__":
    main()
