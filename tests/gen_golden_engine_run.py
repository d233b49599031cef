"""Whole `Engine.run` (Engine.py:3377-3516) of the UNMODIFIED reference engine on the shipped inputs of BASELINE.json
configs 1-3, with fixed seeds: group selection (RandomSelector), move generation (translation / rotation generators),
`transform_coordinates`, the rigid InterMolecularDistanceConstraint pre-filter, the experimental constraints'
compute_before_move / compute_after_move, the engine's Metropolis rule, accept_move / reject_move, scale-factor
refits and shape-function refreshes -- everything the reference does in a step, nothing driven by hand.

    python tests/gen_golden_engine_run.py [--dropin] [--out DIR] [case ...]

Without --dropin (build container) the reference's own compiled kernels run underneath and the result is written to
tests/golden/engine_run_<case>.npz.  With --dropin (GPU box) `fullrmc.Core.pairs_histograms / pairs_distances /
reciprocal_space / atomic_distances / atomic_coordination` are the CUDA drop-in modules of fullrmc_b200.Core
(tests/ref_harness.load_reference(dropin=True)); tests/test_dropin.py requires the two files to be equal array for
array: accepted / tried counts, every constraint's standard error, scale factor, data arrays and the final
coordinates.  The engine object is the harness's array-backed Engine (tests/ref_harness.fake_engine: the real class
with its private fields set directly, because pdbparser / pyrep are not installable offline).
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

_ARGS = sys.argv[1:]
DROPIN = "--dropin" in _ARGS
OUT_DIR = _ARGS[_ARGS.index("--out") + 1] if "--out" in _ARGS else os.path.join(ROOT, "tests", "golden")
ONLY = set(a for i, a in enumerate(_ARGS) if not a.startswith("--") and (i == 0 or _ARGS[i - 1] != "--out"))


def make_engine(fullrmc, arrays):
    box, basis, isPBC, mol, el, elements = arrays
    E = H.fake_engine(fullrmc, box, basis, isPBC, mol, el, elements)
    counts = np.bincount(el, minlength=len(elements))
    object.__setattr__(E, "_Engine__frameOriginalData", {          # what the distance constraint reads (DistanceConstraints.py)
        "_original__elements": list(elements), "_original__allElements": [elements[i] for i in el],
        "_original__elementsIndex": np.ascontiguousarray(el, dtype=np.int32),
        "_original__numberOfAtomsPerElement": {elements[i]: int(counts[i]) for i in range(len(elements))},
        "_original__moleculesIndex": np.ascontiguousarray(mol, dtype=np.int32)})
    return E


def run_case(name, fullrmc, arrays, build, n_steps, seed):
    if ONLY and name not in ONLY:
        return
    from fullrmc.Selectors.RandomSelectors import RandomSelector
    E = make_engine(fullrmc, arrays)
    constraints = build(E)
    for c in constraints:
        H.attach(E, c)
    E.set_group_selector(RandomSelector(E))
    from pyrep import Repository
    object.__setattr__(E, "_Engine__repository", Repository())     # Engine.run insists on one (Engine.py:3406-3408); the stub stores nothing
    random.seed(seed); np.random.seed(seed)
    E.run(numberOfSteps=n_steps, saveFrequency=10 * n_steps, restartPdb=None, ncores=1)
    out = dict(generated=np.int64(E.generated), tried=np.int64(E.tried), accepted=np.int64(E.accepted),
               totalStandardError=np.float64(E.totalStandardError), boxCoordinates=np.asarray(E.boxCoordinates, np.float32).copy(),
               realCoordinates=np.asarray(E.realCoordinates, np.float32).copy(), n_constraints=np.int32(len(constraints)))
    for ci, c in enumerate(constraints):
        out["c%d/class" % ci] = np.array(type(c).__name__)
        out["c%d/standardError" % ci] = np.float64(c.standardError)
        out["c%d/tried" % ci] = np.int64(c.tried); out["c%d/accepted" % ci] = np.int64(c.accepted)
        if hasattr(c, "scaleFactor"):
            out["c%d/scaleFactor" % ci] = np.float32(c.scaleFactor)
        for k, v in c.data.items():
            out["c%d/data_%s" % (ci, k)] = np.asarray(v).copy()
    path = os.path.join(OUT_DIR, "engine_run_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-6s %d steps: generated %d tried %d accepted %d  total standard error %.6f  [%s]  [%d KiB]" % (
        name, n_steps, E.generated, E.tried, E.accepted, float(E.totalStandardError),
        ", ".join("%s %.6g" % (type(c).__name__, float(c.standardError)) for c in constraints), os.path.getsize(path) // 1024))


def main():
    fullrmc = H.load_reference(dropin=DROPIN)
    assert fullrmc is not None, "needs /root/reference (or the package staged by oracle/build_ref.py)"
    if DROPIN:
        import fullrmc_b200
        fullrmc_b200.set_edge_spill(True)        # the reference's unchecked write at bin == histSize
    import gen_golden_constraints as G
    from fullrmc.Globals import FLOAT_TYPE
    from fullrmc.Core.Collection import rebin, convert_Gr_to_gr
    from fullrmc.Core.MoveGenerator import MoveGeneratorCollector
    from fullrmc.Generators.Translations import TranslationGenerator
    from fullrmc.Generators.Rotations import RotationGenerator
    from fullrmc.Constraints.PairDistributionConstraints import PairDistributionConstraint
    from fullrmc.Constraints.PairCorrelationConstraints import PairCorrelationConstraint
    from fullrmc.Constraints.StructureFactorConstraints import ReducedStructureFactorConstraint
    from fullrmc.Constraints.DistanceConstraints import InterMolecularDistanceConstraint
    EX = H.examples_dir()

    # ---- config 1, Examples/atomicNiTi/run.py:41-58, 102-103: G(r) + reduced S(Q) with scale-factor refits every 10
    #      accepted moves, the inter-molecular distance constraint, single-atom groups
    d = os.path.join(EX, "atomicNiTi")
    arrays = G.engine_arrays(*G.read_pdb(os.path.join(d, "system.pdb")))
    def niti(E):
        pdf = PairDistributionConstraint(experimentalData=os.path.join(d, "experimental.gr"), weighting="atomicNumber")
        Sq = np.transpose(rebin(np.loadtxt(os.path.join(d, "experimental.fq")), bin=0.05)).astype(FLOAT_TYPE)
        rsf = ReducedStructureFactorConstraint(experimentalData=Sq, weighting="atomicNumber")
        emd = InterMolecularDistanceConstraint(defaultDistance=2.2, flexible=True)
        pdf.set_adjust_scale_factor((10, 0.8, 1.2)); rsf.set_adjust_scale_factor((10, 0.8, 1.2))
        E.set_groups(None)                                           # set_groups_as_atoms
        return [pdf, rsf, emd]
    run_case("niti", fullrmc, arrays, niti, 300, 101)

    # ---- config 2, Examples/molecularTHF/run.py:52-60, 129-140: g(r) with data weights, inter-molecular distances,
    #      molecule groups moved by a translation + rotation collector
    d2 = os.path.join(EX, "molecularTHF")
    arrays2 = G.engine_arrays(*G.read_pdb(os.path.join(d2, "thf.pdb")))
    def thf(E):
        _, _, _, gr = convert_Gr_to_gr(np.loadtxt(os.path.join(d2, "thf_pdf.exp")), minIndex=[4, 5, 6])
        dw = np.ones(gr.shape[0]); dw[:np.nonzero(gr[:, 1] > 0)[0][0]] = 0
        pcf = PairCorrelationConstraint(experimentalData=gr.astype(FLOAT_TYPE), weighting="atomicNumber", dataWeights=dw)
        emd = InterMolecularDistanceConstraint(defaultDistance=1.5)
        mol = arrays2[3]
        E.set_groups([np.flatnonzero(mol == m).tolist() for m in range(int(mol.max()) + 1)])
        for g in E.groups:
            g.set_move_generator(MoveGeneratorCollector(collection=[TranslationGenerator(amplitude=0.2), RotationGenerator(amplitude=2)],
                                                        randomize=True))
        return [pcf, emd]
    run_case("thf", fullrmc, arrays2, thf, 120, 102)

    # ---- config 3, Examples/SiOxNanosphere/run.py:41-61: non-periodic G(r) with its shape function (refreshed every 40
    #      accepted moves here), element-typed minimum distances, single-atom groups
    d3 = os.path.join(EX, "SiOxNanosphere")
    arrays3 = G.engine_arrays(*G.read_pdb(os.path.join(d3, "SiOx.pdb")))
    def siox(E):
        object.__setattr__(E, "_Engine__numberDensity", FLOAT_TYPE(0.0125))          # run.py:61 set_number_density
        object.__setattr__(E, "_Engine__volume", FLOAT_TYPE(E.numberOfAtoms / 0.0125))
        pdf = PairDistributionConstraint(experimentalData=os.path.join(d3, "SiOx.gr"), weighting="atomicNumber")
        pdf.set_shape_function_parameters({'rmin': 0., 'rmax': None, 'dr': 0.5, 'qmin': 0.0001, 'qmax': 0.6, 'dq': 0.005,
                                           'updateFreq': 40})
        emd = InterMolecularDistanceConstraint()
        E.set_groups(None)
        return [pdf, emd]
    run_case("siox", fullrmc, arrays3, siox, 250, 103)


if __name__ == "__main__":
    main()
