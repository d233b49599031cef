"""Generates tests/golden/atomic_coordination.npz from the reference's own compiled Extensions/atomic_coordination.pyx
(oracle/_ref, built by oracle/build_ref.py from /root/reference).  Run in the build container:

    python tests/gen_golden_atomic_coordination.py

The fixture holds small systems in every geometry (orthorhombic, triclinic, unwrapped fractional coordinates,
non-periodic) with coordination-number definitions laid out the way AtomicCoordinationNumberConstraint lays them out
(Constraints/AtomicCoordinationConstraints.py:376-400: per definition a sorted core list and shell list, per atom the
definitions it is a core of / in the shell of), and the outputs of every public function of the module."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def definitions(rng, n, el, nT, ndef):
    """ndef definitions (core element, shell element, lower, upper) -> the constraint's six lists"""
    cores, shells, lowers, uppers = [], [], [], []
    for d in range(ndef):
        ce, se = int(rng.integers(0, nT)), int(rng.integers(0, nT))
        c = np.nonzero(el == ce)[0]
        s = np.nonzero(el == se)[0]
        if d == ndef - 1:                         # a definition over a sparse hand-picked subset, unsorted lists allowed
            c = rng.permutation(c)[: max(1, len(c) // 3)]
        cores.append(c.astype(np.int32))
        shells.append(s.astype(np.int32))
        lo = np.float32(rng.random() * 1.5)
        lowers.append(lo)
        uppers.append(np.float32(lo + 1.0 + rng.random() * 3.0))
    as_core = [[] for _ in range(n)]
    in_shell = [[] for _ in range(n)]
    for d in range(ndef):
        for i in cores[d]:
            as_core[int(i)].append(d)
        for i in shells[d]:
            in_shell[int(i)].append(d)
    return cores, shells, lowers, uppers, as_core, in_shell


def systems():
    rng = np.random.default_rng(1103)
    tri = np.array([[19, 0, 0], [2.5, 18, 0], [-1.5, 3, 17]], np.float32)
    out = []
    for name, n, nT, basis, pbc, spread, ndef in (("ortho", 700, 3, np.diag([18.0, 19.0, 17.0]).astype(np.float32), True, 0.0, 4),
                                                  ("tri", 640, 2, tri, True, 0.0, 3),
                                                  ("tri_unwrapped", 500, 3, tri, True, 1.3, 5),
                                                  ("non_periodic", 600, 2, np.eye(3, dtype=np.float32), False, 0.0, 3)):
        box = (rng.random((n, 3)) * (1 + 2 * spread) - spread).astype(np.float32)
        if not pbc:
            box = (box * 17.0).astype(np.float32)
        el = rng.integers(0, nT, n).astype(np.int32)
        defs = definitions(rng, n, el, nT, ndef)
        idx = rng.integers(0, n, 9).astype(np.int32)
        out.append((name, box, basis, pbc, el, defs, idx))
    return out


def pack_lists(prefix, lists, out):
    out[prefix + "/n"] = np.int32(len(lists))
    for i, a in enumerate(lists):
        out["%s/%d" % (prefix, i)] = np.asarray(a, dtype=np.int32)


def unpack_lists(g, prefix, as_list=False):
    r = [g["%s/%d" % (prefix, i)] for i in range(int(g[prefix + "/n"]))]
    return [[int(x) for x in a] for a in r] if as_list else r


def run_all(ac, pd, box, basis, pbc, defs, idx):
    """every public function of the module -> dict of outputs"""
    cores, shells, lowers, uppers, as_core, in_shell = defs
    ndef = len(cores)
    lists = dict(coresIndexes=cores, shellsIndexes=shells, lowerShells=list(lowers), upperShells=list(uppers),
                 asCoreDefIdxs=as_core, inShellDefIdxs=in_shell)
    res = {}
    data = np.zeros(ndef, np.float32)
    ac.all_atoms_coord_number_coords(boxCoords=box, basis=basis, isPBC=pbc, coordNumData=data, ncores=1, **lists)
    res["all_coords"] = data.copy()
    data = np.zeros(ndef, np.float32)
    ac.multi_atoms_coord_number_coords(indexes=idx, boxCoords=box, basis=basis, isPBC=pbc, coordNumData=data, ncores=1, **lists)
    res["multi_coords"] = data.copy()
    data = np.full(ndef, 3.0, np.float32)       # accumulates into what is there
    ac.single_atom_coord_number_coords(atomIndex=int(idx[0]), boxCoords=box, basis=basis, isPBC=pbc, coordNumData=data, ncores=1, **lists)
    res["single_coords"] = data.copy()
    res["single_shell_coords"] = np.array([ac.single_atom_single_shell_coords(int(a), shells[d % ndef], box, basis, pbc,
                                                                              lowers[d % ndef], uppers[d % ndef], 1)
                                           for d, a in enumerate(idx)], np.float32)
    res["multi_shells_coords"] = np.asarray(ac.single_atom_multi_shells_coords(int(idx[1]), shells, box, basis, pbc,
                                                                               np.array(lowers, np.float32), np.array(uppers, np.float32), 1))
    # the *_totdists forms on the distance rows of the same atoms
    rows = [np.asarray(pd.pairs_distances_to_indexcoords(atomIndex=int(a), coords=box, basis=basis, isPBC=pbc, allAtoms=True, ncores=1))
            for a in idx]
    data = np.zeros(ndef, np.float32)
    ac.multi_atoms_coord_number_totdists(indexes=idx, distances=rows, coordNumData=data, ncores=1, **lists)
    res["multi_totdists"] = data.copy()
    data = np.zeros(ndef, np.float32)
    ac.single_atom_coord_number_totdists(atomIndex=int(idx[2]), distances=rows[2], coordNumData=data, ncores=1, **lists)
    res["single_totdists"] = data.copy()
    res["single_shell_totdists"] = np.float32(ac.single_atom_single_shell_totdists(rows[0], shells[0], lowers[0], uppers[0], 1))
    res["single_shell_subdists"] = np.float32(ac.single_atom_single_shell_subdists(rows[0][shells[0]], lowers[0], uppers[0], 1))
    res["multi_shells_totdists"] = np.asarray(ac.single_atom_multi_shells_totdists(rows[1], shells, np.array(lowers, np.float32),
                                                                                   np.array(uppers, np.float32), 1))
    # all_atoms_coord_number_totdists (:317-345) cannot run in the reference: it hands its 2-d ndarray to
    # multi_atoms_coord_number_totdists, whose `distances` is typed `list` -> TypeError on every call
    try:
        ac.all_atoms_coord_number_totdists(distances=np.stack(rows), coordNumData=np.zeros(ndef, np.float32), ncores=1, **lists)
        res["all_totdists_raises_TypeError"] = np.bool_(False)
    except TypeError:
        res["all_totdists_raises_TypeError"] = np.bool_(True)
    return res


def main():
    assert build_ref.build(), "cannot build oracle/_ref"
    build_ref.load()
    ac = importlib.import_module("fullrmc.Core.atomic_coordination")
    pd = importlib.import_module("fullrmc.Core.pairs_distances")
    out = {"names": np.array([s[0] for s in systems()])}
    for name, box, basis, pbc, el, defs, idx in systems():
        cores, shells, lowers, uppers, as_core, in_shell = defs
        out.update({name + "/boxCoords": box, name + "/basis": basis, name + "/isPBC": np.bool_(pbc), name + "/indexes": idx,
                    name + "/lowerShells": np.array(lowers, np.float32), name + "/upperShells": np.array(uppers, np.float32)})
        pack_lists(name + "/cores", cores, out)
        pack_lists(name + "/shells", shells, out)
        pack_lists(name + "/asCore", as_core, out)
        pack_lists(name + "/inShell", in_shell, out)
        for key, val in run_all(ac, pd, box, basis, pbc, defs, idx).items():
            out["%s/out/%s" % (name, key)] = val
    path = os.path.join(ROOT, "tests", "golden", "atomic_coordination.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
