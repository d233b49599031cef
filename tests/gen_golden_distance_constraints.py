"""Generates tests/golden/distance_constraints_<case>.npz by running the UNMODIFIED reference
InterMolecularDistanceConstraint / IntraMolecularDistanceConstraint (Constraints/DistanceConstraints.py) with the
reference's own compiled atomic_distances kernels on the shipped THF and SiOx inputs.

Run in the build container:   python tests/gen_golden_distance_constraints.py

Per case: the engine arrays, what the constraint derived (types, limit arrays, flags) and a trajectory driven like
Engine.__on_runtime_step_try_move: per step the moved group, the moved coordinates, data after the move,
afterMoveStandardError, should_step_get_rejected and the decision."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402
from gen_golden_constraints import read_pdb, engine_arrays, EX  # noqa: E402


def run_case(name, fullrmc, arrays, make_constraint, groups, n_steps, seed, sigma, out_dir, pairs=None):
    box, basis, isPBC, mol, el, elements = arrays
    E = H.fake_engine(fullrmc, box, basis, isPBC, mol, el, elements)
    counts = np.bincount(el, minlength=len(elements))
    allElements = [elements[i] for i in el]
    object.__setattr__(E, "_Engine__frameOriginalData", {
        "_original__elements": list(elements), "_original__allElements": allElements,
        "_original__elementsIndex": np.ascontiguousarray(el, dtype=np.int32),
        "_original__numberOfAtomsPerElement": {elements[i]: int(counts[i]) for i in range(len(elements))},
        "_original__moleculesIndex": np.ascontiguousarray(mol, dtype=np.int32)})
    c = make_constraint(E)
    H.attach(E, c)
    if pairs is not None:
        c.set_pairs_distance(pairs)                               # Examples/SiOxNanosphere/run.py:50
    out = dict(boxCoords=box.copy(), basis=basis, isPBC=np.bool_(isPBC), moleculeIndex=mol, elementIndex=el,
               elements=np.array(elements), typesIndex=np.asarray(c.typesIndex, np.int32), numberOfTypes=np.int32(c.numberOfTypes),
               lowerLimitArray=np.asarray(c.lowerLimitArray, np.float32), upperLimitArray=np.asarray(c.upperLimitArray, np.float32),
               typePairsIndex=np.asarray(c.typePairsIndex, np.int32), interMolecular=np.bool_(c._interMolecular),
               flexible=np.bool_(c.flexible))
    data, err = c.compute_data()
    out["start_number"], out["start_distanceSum"], out["start_stdErr"] = data["number"].copy(), data["distanceSum"].copy(), np.float32(err)
    rng = np.random.default_rng(seed)
    rbasis = np.linalg.inv(basis.astype(np.float64)) if isPBC else np.eye(3)
    logs = dict(idx=[], k=[], moved=[], stdErr=[], rejected=[], accepted=[], number=[], distanceSum=[])
    for step in range(n_steps):
        idx = np.asarray(groups[int(rng.integers(0, len(groups)))], dtype=np.int32)
        shift = (rng.normal(0.0, sigma, (1, 3)) @ rbasis).astype(np.float32)
        moved = (E.boxCoordinates[idx] + shift).astype(np.float32)
        c.compute_before_move(realIndexes=idx, relativeIndexes=idx)
        c.compute_after_move(realIndexes=idx, relativeIndexes=idx, movedBoxCoordinates=moved)
        rejected = bool(c.should_step_get_rejected(c.afterMoveStandardError))
        accept = (not rejected) or step % 4 == 3                     # also exercise accepts the rigid rule would refuse
        logs["stdErr"].append(np.float32(c.afterMoveStandardError)); logs["rejected"].append(rejected); logs["accepted"].append(accept)
        (c.accept_move if accept else c.reject_move)(realIndexes=idx, relativeIndexes=idx)
        if accept:
            E.boxCoordinates[idx] = moved
        logs["idx"].append(np.pad(idx, (0, 64 - idx.shape[0]), constant_values=-1)); logs["k"].append(idx.shape[0])
        logs["moved"].append(np.pad(moved, ((0, 64 - idx.shape[0]), (0, 0))))
        logs["number"].append(c.data["number"].copy()); logs["distanceSum"].append(c.data["distanceSum"].copy())
    out["steps/idx"] = np.array(logs["idx"], np.int32); out["steps/k"] = np.array(logs["k"], np.int32)
    out["steps/moved"] = np.array(logs["moved"], np.float32); out["steps/stdErr_after"] = np.array(logs["stdErr"], np.float32)
    out["steps/rejected"] = np.array(logs["rejected"], np.bool_); out["steps/accepted"] = np.array(logs["accepted"], np.bool_)
    out["steps/number"] = np.array(logs["number"], np.int32); out["steps/distanceSum"] = np.array(logs["distanceSum"], np.float32)
    out["final_stdErr"] = np.float32(c.standardError)
    out["final_value"] = np.asarray(c._get_constraint_value(), np.float32)
    path = os.path.join(out_dir, "distance_constraints_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-12s %5d atoms, %d steps (%d accepted, %d flagged for rejection), stdErr %s -> %s  [%d KiB]" % (
        name, box.shape[0], n_steps, int(np.sum(logs["accepted"])), int(np.sum(logs["rejected"])), float(err), float(c.standardError),
        os.path.getsize(path) // 1024))


def main():
    fullrmc = H.load_reference()
    assert fullrmc is not None, "needs /root/reference"
    from fullrmc.Constraints.DistanceConstraints import InterMolecularDistanceConstraint, IntraMolecularDistanceConstraint
    out_dir = os.path.join(ROOT, "tests", "golden")
    # THF (Examples/molecularTHF/run.py): inter-molecular minimum distances, molecule moves
    arrays = engine_arrays(*read_pdb(os.path.join(EX, "molecularTHF", "thf.pdb")))
    mol = arrays[3]
    groups = [np.flatnonzero(mol == m).tolist() for m in range(int(mol.max()) + 1)]
    run_case("thf_inter", fullrmc, arrays, lambda E: InterMolecularDistanceConstraint(defaultDistance=2.2, flexible=False),
             groups, 30, 5, 0.6, out_dir)
    # the intra-molecular class on the same system, element types, single-atom moves, flexible
    n = arrays[0].shape[0]
    run_case("thf_intra", fullrmc, arrays, lambda E: IntraMolecularDistanceConstraint(defaultDistance=1.8, typeDefinition="element"),
             [[i] for i in range(n)], 30, 6, 0.3, out_dir)
    # SiOx nanosphere, non-periodic (Examples/SiOxNanosphere/run.py:48-51)
    arrays = engine_arrays(*read_pdb(os.path.join(EX, "SiOxNanosphere", "SiOx.pdb")))
    n = arrays[0].shape[0]
    run_case("siox_inter", fullrmc, arrays, lambda E: InterMolecularDistanceConstraint(), [[i] for i in range(n)], 30, 7, 1.5, out_dir,
             pairs=[('si', 'si', 1.75), ('o', 'o', 1.10), ('si', 'o', 1.30)])


if __name__ == "__main__":
    main()
