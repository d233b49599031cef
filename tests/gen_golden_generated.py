"""Generates tests/golden/generated_<case>.npz: whole `Engine.run`s of the UNMODIFIED reference engine whose random
numbers come from the counter-based plug-ins of fullrmc_b200/engine_plugins.py (group selector, translation generator,
acceptance number -- the engine's own extension points), with the reference's own compiled kernels underneath.

    python tests/gen_golden_generated.py [case ...]          (build container)

tests/test_generated_runs.py gives the device the same seed and first counter (`DeviceStore.run_generated`: selection,
translation, transform_coordinates, evaluation, decision and move application all on the device) and requires the same
accepted count, standard errors, data arrays, real and box coordinates -- the device half of the random-number contract
of fullrmc_b200/rng.py, SURVEY section 8f rank 2.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

ONLY = set(a for a in sys.argv[1:] if not a.startswith("--"))


def run_case(name, fullrmc, G, E_maker, arrays, build, groups, n_steps, seed, first_counter, amplitude, out_dir):
    if ONLY and name not in ONLY:
        return
    from fullrmc_b200 import engine_plugins
    E = E_maker(fullrmc, arrays)
    constraints = build(E)
    kinds = [k for _, k in constraints]
    for c, _ in constraints:
        H.attach(E, c)
    E.set_groups(groups)
    stream = engine_plugins.install(fullrmc, E, seed, first_counter, amplitude)
    from pyrep import Repository
    object.__setattr__(E, "_Engine__repository", Repository())
    box, basis, isPBC, mol, el, elements = arrays
    out = dict(basis=basis, isPBC=np.bool_(isPBC), elements=np.array(elements), volume=np.float32(E.volume),
               numberDensity=np.float32(E.numberDensity), n_constraints=np.int32(len(constraints)),
               boxCoords=box.copy(), realCoords=np.asarray(E.realCoordinates, np.float32).copy(),
               reciprocalBasis=np.asarray(E.reciprocalBasisVectors, np.float32).copy(), moleculeIndex=mol, elementIndex=el,
               seed=np.uint64(seed), first_counter=np.uint64(first_counter), amplitude=np.float32(amplitude), n_steps=np.int32(n_steps),
               group_offsets=np.cumsum([0] + [len(g) for g in (groups if groups is not None else [[i] for i in range(box.shape[0])])]).astype(np.int32),
               group_indexes=np.concatenate([np.asarray(g, np.int32) for g in (groups if groups is not None else [[i] for i in range(box.shape[0])])]))
    for ci, (c, kind) in enumerate(constraints):
        for k, v in G.describe(c, kind).items():
            out["c%d/%s" % (ci, k)] = v
    E.run(numberOfSteps=n_steps, saveFrequency=10 * n_steps, restartPdb=None, ncores=1)
    assert stream.next_counter == first_counter + n_steps
    out.update(generated=np.int64(E.generated), tried=np.int64(E.tried), accepted=np.int64(E.accepted),
               totalStandardError=np.float32(E.totalStandardError),
               final_boxCoords=np.asarray(E.boxCoordinates, np.float32).copy(),
               final_realCoords=np.asarray(E.realCoordinates, np.float32).copy())
    for ci, (c, kind) in enumerate(constraints):
        out["c%d/final_stdErr" % ci] = np.float32(c.standardError)
        out["c%d/final_intra" % ci], out["c%d/final_inter" % ci] = c.data["intra"].copy(), c.data["inter"].copy()
        out["c%d/final_scaleFactor" % ci] = np.float32(c.scaleFactor)
    path = os.path.join(out_dir, "generated_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-8s %d steps: tried %d accepted %d  total standard error %.6f  [%d KiB]" % (
        name, n_steps, E.tried, E.accepted, float(E.totalStandardError), os.path.getsize(path) // 1024))


def main():
    fullrmc = H.load_reference()
    assert fullrmc is not None, "needs /root/reference (or the package staged by oracle/build_ref.py)"
    sys.argv = sys.argv[:1]
    import gen_golden_constraints as G
    import gen_golden_engine_run as R
    from fullrmc.Globals import FLOAT_TYPE
    from fullrmc.Core.Collection import rebin, convert_Gr_to_gr
    from fullrmc.Constraints.PairDistributionConstraints import PairDistributionConstraint
    from fullrmc.Constraints.PairCorrelationConstraints import PairCorrelationConstraint
    from fullrmc.Constraints.StructureFactorConstraints import StructureFactorConstraint, ReducedStructureFactorConstraint
    out_dir = os.path.join(ROOT, "tests", "golden")
    EX = H.examples_dir()

    # periodic, atomic groups: Examples/atomicNiTi (PDF + reduced S(Q); no refit: the batch kernel's restriction)
    d = os.path.join(EX, "atomicNiTi")
    arrays = G.engine_arrays(*G.read_pdb(os.path.join(d, "system.pdb")))
    def niti(E):
        pdf = PairDistributionConstraint(experimentalData=os.path.join(d, "experimental.gr"), weighting="atomicNumber")
        Sq = np.transpose(rebin(np.loadtxt(os.path.join(d, "experimental.fq")), bin=0.05)).astype(FLOAT_TYPE)
        rsf = ReducedStructureFactorConstraint(experimentalData=Sq, weighting="atomicNumber")
        return [(pdf, "PDF"), (rsf, "RSQ")]
    run_case("niti", fullrmc, G, R.make_engine, arrays, niti, None, 400, 0x1234ABCD5678EF01, 1000, 0.3, out_dir)

    # the same as shipped (Examples/atomicNiTi/run.py:102-103): both constraints refit their scale factor every 10 accepted moves
    def niti_sf(E):
        cons = niti(E)
        for c, _ in cons:
            c.set_adjust_scale_factor((10, 0.8, 1.2))
        return cons
    run_case("niti_sf", fullrmc, G, R.make_engine, arrays, niti_sf, None, 400, 2024, 0, 0.3, out_dir)

    # periodic, molecule groups (k = 13): Examples/molecularTHF, g(r) with data weights
    d2 = os.path.join(EX, "molecularTHF")
    arrays2 = G.engine_arrays(*G.read_pdb(os.path.join(d2, "thf.pdb")))
    def thf(E):
        _, _, _, gr = convert_Gr_to_gr(np.loadtxt(os.path.join(d2, "thf_pdf.exp")), minIndex=[4, 5, 6])
        dw = np.ones(gr.shape[0]); dw[:np.nonzero(gr[:, 1] > 0)[0][0]] = 0
        return [(PairCorrelationConstraint(experimentalData=gr.astype(FLOAT_TYPE), weighting="atomicNumber", dataWeights=dw), "PCF")]
    mol = arrays2[3]
    groups2 = [np.flatnonzero(mol == m).tolist() for m in range(int(mol.max()) + 1)]
    run_case("thf", fullrmc, G, R.make_engine, arrays2, thf, groups2, 150, 77, 0, 0.2, out_dir)

    # non-periodic, atomic groups: Examples/SiOxNanosphere
    d3 = os.path.join(EX, "SiOxNanosphere")
    arrays3 = G.engine_arrays(*G.read_pdb(os.path.join(d3, "SiOx.pdb")))
    def siox(E):
        object.__setattr__(E, "_Engine__numberDensity", FLOAT_TYPE(0.0125))
        object.__setattr__(E, "_Engine__volume", FLOAT_TYPE(E.numberOfAtoms / 0.0125))
        return [(PairDistributionConstraint(experimentalData=os.path.join(d3, "SiOx.gr"), weighting="atomicNumber"), "PDF")]
    run_case("siox", fullrmc, G, R.make_engine, arrays3, siox, None, 300, 3, 5, (0.05, 0.25), out_dir)

    # synthetic triclinic, few atoms and many steps: the same atom is moved again and again (conflicts inside a launch),
    # tolerance 0, molecules of 3 moved as groups, G(r) + full S(Q)
    rng = np.random.default_rng(47)
    n = 600
    box = rng.random((n, 3), dtype=np.float32)
    basis = np.array([[21, 0, 0], [3, 20, 0], [-2, 4, 19.5]], dtype=np.float32)
    el = rng.integers(0, 3, n).astype(np.int32)
    mol4 = (np.arange(n) // 3).astype(np.int32)
    arrays4 = (box, basis, True, mol4, el, ["o", "si", "ti"])
    def synth(E):
        r = (0.05 + 0.05 * np.arange(180)).astype(np.float32)
        pdf = PairDistributionConstraint(experimentalData=np.stack([r, rng.normal(0, 0.2, 180).astype(np.float32)], 1).astype(np.float32),
                                         weighting="atomicNumber", scaleFactor=0.95)
        q = np.linspace(0.6, 14.0, 120).astype(np.float32)
        sf = StructureFactorConstraint(experimentalData=np.stack([q, 1 + rng.normal(0, 0.1, 120).astype(np.float32)], 1).astype(np.float32),
                                       weighting="atomicNumber", scaleFactor=1.05, rmax=9.0)
        return [(pdf, "PDF"), (sf, "SQ")]
    groups4 = [[3 * m, 3 * m + 1, 3 * m + 2] for m in range(n // 3)]
    run_case("synth", fullrmc, G, R.make_engine, arrays4, synth, groups4, 600, 99, 123456789012, 0.4, out_dir)


if __name__ == "__main__":
    main()
