"""Generates tests/golden/removal_<case>.npz: trajectories that mix moves with ATOM REMOVALS, recorded from the
UNMODIFIED reference constraint classes and the reference Engine's own bookkeeping
(compute_as_if_amputated / accept_amputation / reject_amputation of PairDistributionConstraint,
PairCorrelationConstraint, StructureFactorConstraint; Engine._on_collector_collect_atom, Engine.py:758-797), driven like
Engine.__on_runtime_step_try_remove / __on_runtime_step_try_move (Engine.py:3231-3338).  SURVEY section 8f rank 4.

Run in the build container:   python tests/gen_golden_removal.py [case ...]

tests/test_removal.py replays them through the device store (frmc_propose_amputation / frmc_accept_amputation /
frmc_model_set_constants): chi^2 of every step, the weighting schemes, the final data arrays and totals bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

ONLY = set(a for a in sys.argv[1:] if not a.startswith("--"))


def run_case(name, fullrmc, G, arrays, make_constraints, n_steps, seed, sigma, remove_every, allow_fit, out_dir):
    if ONLY and name not in ONLY:
        return
    box, basis, isPBC, mol, el, elements = arrays
    E = H.fake_engine(fullrmc, box, basis, isPBC, mol, el, elements)
    n = box.shape[0]
    # what Engine._on_collector_collect_atom touches beyond the constraints' needs (Engine.py:765-793)
    names = [elements[i] for i in el]
    priv = dict(moleculesName=["m%d" % m for m in mol], namesIndex=np.array(el, dtype=np.int32), allNames=list(names),
                names=sorted(set(names)), numberOfAtomsPerName={e: names.count(e) for e in set(names)},
                numberOfMolecules=len(set(mol.tolist())))
    for k, v in priv.items():
        object.__setattr__(E, "_Engine__" + k, v)
    E._atomsCollector.set_data_keys(["realCoordinates", "boxCoordinates", "moleculesIndex", "moleculesName", "elementsIndex",
                                     "allElements", "namesIndex", "allNames"])
    object.__setattr__(E, "_RT_moveGenerator", type("RemoveGeneratorStandIn", (object,), {"allowFittingScaleFactor": bool(allow_fit)})())
    constraints = make_constraints(E)
    out = dict(basis=basis, isPBC=np.bool_(isPBC), elements=np.array(elements), volume=np.float32(E.volume),
               numberDensity=np.float32(E.numberDensity), n_constraints=np.int32(len(constraints)), allow_fit=np.bool_(allow_fit),
               boxCoords=box.copy(), moleculeIndex=mol, elementIndex=el)
    for ci, (c, kind) in enumerate(constraints):
        H.attach(E, c)
        for k, v in G.describe(c, kind).items():
            out["c%d/%s" % (ci, k)] = v
        out["c%d/elementsWeight" % ci] = np.array([c._elementsWeight[e] for e in elements], np.float64)
    start = []
    for ci, (c, kind) in enumerate(constraints):
        c.compute_data()
        start.append(np.float32(c.standardError))
    out["start_stdErr"] = np.array(start, np.float32)
    rng = np.random.default_rng(seed)
    rbasis = np.linalg.inv(basis.astype(np.float64)) if isPBC else np.eye(3)
    alive = list(range(n))                       # real index of every remaining atom, in relative order
    kind_log, idx_log, moved_log, chi_log, acc_log, sf_log, w_log, rho_log = [], [], [], [], [], [], [], []
    total_old = sum(float(c.standardError) for c, _ in constraints)
    for step in range(n_steps):
        removal = remove_every and step % remove_every == remove_every - 1
        rel = int(rng.integers(0, len(alive)))
        relIdx = np.array([rel], dtype=np.int32)
        realIdx = np.array([alive[rel]], dtype=np.int32)
        if removal:
            for c, _ in constraints:
                c.compute_as_if_amputated(realIndex=realIdx, relativeIndex=relIdx)
            chis = [np.float32(c.amputationStandardError) for c, _ in constraints]
            moved = np.zeros((1, 3), np.float32)
        else:
            shift = (rng.normal(0.0, sigma, (1, 3)) @ rbasis).astype(np.float32)
            moved = (E.boxCoordinates[relIdx] + shift).astype(np.float32)
            for c, _ in constraints:
                c.compute_before_move(realIndexes=realIdx, relativeIndexes=relIdx)
                c.compute_after_move(realIndexes=realIdx, relativeIndexes=relIdx, movedBoxCoordinates=moved)
            chis = [np.float32(c.afterMoveStandardError) for c, _ in constraints]
        sf_log.append([np.float32(c._fittedScaleFactor) for c, _ in constraints])
        total_new = sum(float(x) for x in chis)
        accept = total_new <= total_old or step % 4 == 3            # also uphill accepts; removals are accepted and refused
        if removal and step % (4 * remove_every) == remove_every - 1:
            accept = False                                           # make sure some removals are refused
        if removal:
            for c, _ in constraints:
                (c.accept_amputation if accept else c.reject_amputation)(realIndex=realIdx, relativeIndex=relIdx)
            if accept:
                E._on_collector_collect_atom(realIndex=int(realIdx[0]))      # Engine.py:3266
                alive.pop(rel)
        else:
            for c, _ in constraints:
                (c.accept_move if accept else c.reject_move)(realIndexes=realIdx, relativeIndexes=relIdx)
            if accept:
                E.boxCoordinates[relIdx] = moved
        if accept:
            total_old = total_new
            object.__setattr__(E, "_Engine__accepted", E.accepted + 1)
        kind_log.append(1 if removal else 0); idx_log.append(rel); moved_log.append(moved[0]); chi_log.append(chis)
        acc_log.append(accept)
        ws = []
        for c, kind in constraints:                                   # the weighting scheme every constraint holds now
            pre = "_StructureFactorConstraint__" if kind in ("SQ", "RSQ") else "_PairDistributionConstraint__"
            w = c.weightingScheme if kind == "PCF" and hasattr(c, "weightingScheme") else getattr(c, pre + "weightingScheme")
            pairs = getattr(c, pre + "elementsPairs")
            ws.append([w.get("%s-%s" % p, w.get("%s-%s" % (p[1], p[0]))) for p in pairs])
        w_log.append(ws)
        rho_log.append(np.float32(E.numberDensity))
    out["steps/kind"] = np.array(kind_log, np.int32)
    out["steps/idx"] = np.array(idx_log, np.int32)
    out["steps/moved"] = np.array(moved_log, np.float32)
    out["steps/chi2_after"] = np.array(chi_log, np.float32)
    out["steps/accepted"] = np.array(acc_log, np.bool_)
    out["steps/scale_used"] = np.array(sf_log, np.float32)
    out["steps/pair_w"] = np.array(w_log, np.float32)
    out["steps/numberDensity"] = np.array(rho_log, np.float32)
    for ci, (c, kind) in enumerate(constraints):
        out["c%d/final_intra" % ci], out["c%d/final_inter" % ci] = c.data["intra"].copy(), c.data["inter"].copy()
        out["c%d/final_stdErr" % ci] = np.float32(c.standardError)
        out["c%d/final_scaleFactor" % ci] = np.float32(c.scaleFactor)
        out["c%d/final_total" % ci] = G.fit_total(c, kind, E)
        # compute_data from scratch on the final configuration: what a store rebuilt after the removals must give
        d2, e2 = c.compute_data(update=False)
        out["c%d/recomputed_intra" % ci], out["c%d/recomputed_inter" % ci] = d2["intra"].copy(), d2["inter"].copy()
    out["final_boxCoords"] = np.asarray(E.boxCoordinates, np.float32).copy()
    out["final_alive"] = np.array(alive, np.int32)
    path = os.path.join(out_dir, "removal_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-14s %5d atoms -> %d, %d constraints, %d steps (%d removals tried, %d accepted; %d moves accepted)   [%d KiB]" % (
        name, n, len(alive), len(constraints), n_steps, int(np.sum(kind_log)),
        int(np.sum(np.array(acc_log) & (np.array(kind_log) == 1))), int(np.sum(np.array(acc_log) & (np.array(kind_log) == 0))),
        os.path.getsize(path) // 1024))


def main():
    fullrmc = H.load_reference()
    assert fullrmc is not None, "needs /root/reference (or the package staged by oracle/build_ref.py)"
    sys.argv = sys.argv[:1]                      # gen_golden_constraints parses the command line at import
    import gen_golden_constraints as G
    from fullrmc.Globals import FLOAT_TYPE
    from fullrmc.Core.Collection import rebin
    from fullrmc.Constraints.PairDistributionConstraints import PairDistributionConstraint
    from fullrmc.Constraints.PairCorrelationConstraints import PairCorrelationConstraint
    from fullrmc.Constraints.StructureFactorConstraints import StructureFactorConstraint, ReducedStructureFactorConstraint
    out_dir = os.path.join(ROOT, "tests", "golden")
    EX = H.examples_dir()

    # periodic, two constraints on two r-grids: Examples/atomicNiTi (PDF + reduced S(Q))
    d = os.path.join(EX, "atomicNiTi")
    arrays = G.engine_arrays(*G.read_pdb(os.path.join(d, "system.pdb")))
    def niti(E):
        pdf = PairDistributionConstraint(experimentalData=os.path.join(d, "experimental.gr"), weighting="atomicNumber")
        Sq = np.transpose(rebin(np.loadtxt(os.path.join(d, "experimental.fq")), bin=0.05)).astype(FLOAT_TYPE)
        rsf = ReducedStructureFactorConstraint(experimentalData=Sq, weighting="atomicNumber")
        return [(pdf, "PDF"), (rsf, "RSQ")]
    run_case("niti", fullrmc, G, arrays, niti, 36, 21, 0.15, 3, False, out_dir)

    # the same with the scale-factor refit schedule on and allowFittingScaleFactor=True
    def niti_sf(E):
        cons = niti(E)
        for c, _ in cons:
            c.set_adjust_scale_factor((4, 0.8, 1.2))
        return cons
    run_case("niti_fit", fullrmc, G, arrays, niti_sf, 30, 22, 0.15, 3, True, out_dir)

    # non-periodic (the number density does not follow the removals, Engine.py:796): Examples/SiOxNanosphere
    d3 = os.path.join(EX, "SiOxNanosphere")
    arrays3 = G.engine_arrays(*G.read_pdb(os.path.join(d3, "SiOx.pdb")))
    def siox(E):
        object.__setattr__(E, "_Engine__numberDensity", FLOAT_TYPE(0.0125))
        object.__setattr__(E, "_Engine__volume", FLOAT_TYPE(E.numberOfAtoms / 0.0125))
        return [(PairDistributionConstraint(experimentalData=os.path.join(d3, "SiOx.gr"), weighting="atomicNumber"), "PDF")]
    run_case("siox", fullrmc, G, arrays3, siox, 30, 23, 0.2, 2, False, out_dir)

    # synthetic triclinic, 4 elements, molecules of 3: g(r) with data weights + full S(Q) with scale factors
    rng = np.random.default_rng(46)
    n = 2400
    box = (rng.random((n, 3), dtype=np.float32) * np.float32(1.6) - np.float32(0.3)).astype(np.float32)
    basis = np.array([[33, 0, 0], [5, 32, 0], [-3.5, 6.5, 31]], dtype=np.float32)
    el = rng.integers(0, 4, n).astype(np.int32)
    mol = (np.arange(n) // 3).astype(np.int32)
    arrays4 = (box, basis, True, mol, el, ["o", "si", "ti", "zr"])
    def synth(E):
        r = (0.05 + 0.05 * np.arange(300)).astype(np.float32)
        pcf = PairCorrelationConstraint(experimentalData=np.stack([r, 1 + rng.normal(0, 0.2, 300).astype(np.float32)], 1).astype(np.float32),
                                        weighting="atomicNumber", scaleFactor=0.97, dataWeights=rng.random(300))
        q = np.linspace(0.6, 14.0, 150).astype(np.float32)
        sf = StructureFactorConstraint(experimentalData=np.stack([q, 1 + rng.normal(0, 0.1, 150).astype(np.float32)], 1).astype(np.float32),
                                       weighting="atomicNumber", scaleFactor=1.05)
        return [(pcf, "PCF"), (sf, "SQ")]
    run_case("synth", fullrmc, G, arrays4, synth, 30, 24, 0.25, 2, False, out_dir)


if __name__ == "__main__":
    main()
