"""Generates tests/golden/coordination_constraint_<case>.npz by running the UNMODIFIED reference
AtomicCoordinationNumberConstraint (Constraints/AtomicCoordinationConstraints.py) with the reference's own compiled
atomic_coordination kernels on the shipped SiOx (non-periodic) and NiTi (periodic) inputs.

Run in the build container:   python tests/gen_golden_coordination_constraint.py

Per case: the engine arrays, the definition lists the constraint derived, and a trajectory driven like
Engine.__on_runtime_step_try_move: per step the moved atom, its new coordinates, afterMoveStandardError, the decision and
the data after it."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402
from gen_golden_constraints import read_pdb, engine_arrays, EX  # noqa: E402
from gen_golden_atomic_coordination import pack_lists  # noqa: E402


def pack_csr(prefix, lists, out):
    """per-atom lists as offsets + values (thousands of tiny arrays would bloat the archive)"""
    out[prefix + "/offsets"] = np.concatenate([[0], np.cumsum([len(a) for a in lists])]).astype(np.int64)
    out[prefix + "/values"] = (np.concatenate([np.asarray(a, np.int32) for a in lists]) if len(lists) else np.zeros(0, np.int32)).astype(np.int32)


def unpack_csr(g, prefix):
    off, val = g[prefix + "/offsets"], g[prefix + "/values"]
    return [val[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def run_case(name, fullrmc, arrays, definition, n_steps, seed, sigma, out_dir):
    from fullrmc.Constraints.AtomicCoordinationConstraints import AtomicCoordinationNumberConstraint
    box, basis, isPBC, mol, el, elements = arrays
    E = H.fake_engine(fullrmc, box, basis, isPBC, mol, el, elements)
    allElements = [elements[i] for i in el]
    object.__setattr__(E, "_Engine__frameOriginalData", {
        "_original__elements": list(elements), "_original__allElements": allElements, "_original__names": list(elements),
        "_original__allNames": list(allElements), "_original__numberOfAtoms": int(box.shape[0])})
    c = AtomicCoordinationNumberConstraint()
    H.attach(E, c)
    c.set_coordination_number_definition(definition)
    out = dict(boxCoords=box.copy(), basis=basis, isPBC=np.bool_(isPBC), lowerShells=np.array(c.lowerShells, np.float32),
               upperShells=np.array(c.upperShells, np.float32), minAtoms=np.array(c.minAtoms, np.float32),
               maxAtoms=np.array(c.maxAtoms, np.float32), weights=np.asarray(c.weights, np.float32))
    pack_lists("cores", c.coresIndexes, out)
    pack_lists("shells", c.shellsIndexes, out)
    pack_csr("asCore", c.asCoreDefIdxs, out)
    pack_csr("inShell", c.inShellDefIdxs, out)
    data, err = c.compute_data()
    out["start_data"], out["start_stdErr"] = np.asarray(data, np.float32).copy(), np.float32(err)
    rng = np.random.default_rng(seed)
    rbasis = np.linalg.inv(basis.astype(np.float64)) if isPBC else np.eye(3)
    logs = dict(idx=[], moved=[], stdErr=[], accepted=[], data=[])
    n = box.shape[0]
    for step in range(n_steps):
        idx = np.array([int(rng.integers(0, n))], dtype=np.int32)
        moved = (E.boxCoordinates[idx] + (rng.normal(0.0, sigma, (1, 3)) @ rbasis)).astype(np.float32)
        c.compute_before_move(realIndexes=idx, relativeIndexes=idx)
        c.compute_after_move(realIndexes=idx, relativeIndexes=idx, movedBoxCoordinates=moved)
        after = c.afterMoveStandardError
        accept = step % 3 != 1                                    # a third of the moves is refused whatever they do
        logs["stdErr"].append(np.float32(after)); logs["accepted"].append(accept)
        (c.accept_move if accept else c.reject_move)(realIndexes=idx, relativeIndexes=idx)
        if accept:
            E.boxCoordinates[idx] = moved
        logs["idx"].append(idx[0]); logs["moved"].append(moved[0]); logs["data"].append(np.asarray(c.data, np.float32).copy())
    out["steps/idx"] = np.array(logs["idx"], np.int32); out["steps/moved"] = np.array(logs["moved"], np.float32)
    out["steps/stdErr_after"] = np.array(logs["stdErr"], np.float32); out["steps/accepted"] = np.array(logs["accepted"], np.bool_)
    out["steps/data"] = np.array(logs["data"], np.float32)
    out["final_stdErr"] = np.float32(c.standardError)
    recount, err2 = c.compute_data(update=False)
    out["final_recount"] = np.asarray(recount, np.float32)
    path = os.path.join(out_dir, "coordination_constraint_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-10s %5d atoms, %d steps (%d accepted), stdErr %s -> %s, data %s [%d KiB]" % (
        name, n, n_steps, int(np.sum(logs["accepted"])), float(err), float(c.standardError), np.asarray(c.data), os.path.getsize(path) // 1024))


def main():
    fullrmc = H.load_reference()
    assert fullrmc is not None, "needs /root/reference"
    out_dir = os.path.join(ROOT, "tests", "golden")
    # SiOx nanosphere, non-periodic: Si-O first shell and O-Si, Si-Si second shell with a weight
    arrays = engine_arrays(*read_pdb(os.path.join(EX, "SiOxNanosphere", "SiOx.pdb")))
    els = list(arrays[5])
    si, o = [e for e in els if e.lower() == "si"][0], [e for e in els if e.lower() == "o"][0]
    run_case("siox", fullrmc, arrays, [(si, o, 1.0, 3.5, 3.5, 4.5), (o, si, 1.0, 3.5, 2, 2), (si, si, 2.0, 5.0, 6, 10, 2.0)], 40, 11, 0.8, out_dir)
    # NiTi, periodic: element shells plus a definition given by explicit atom indexes
    arrays = engine_arrays(*read_pdb(os.path.join(EX, "atomicNiTi", "system.pdb")))
    els = list(arrays[5])
    run_case("niti", fullrmc, arrays, [(els[0], els[1], 2.0, 3.0, 8.5, 9), (els[1], els[1], 2.5, 3.3, 4, 5.5, 0.5),
                                       (list(range(0, 600, 3)), els[0], 0.0, 3.1, 7.5, 9, 3.0)], 40, 12, 0.3, out_dir)


if __name__ == "__main__":
    main()
