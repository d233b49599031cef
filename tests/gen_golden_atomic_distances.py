"""Generates tests/golden/atomic_distances.npz from the reference's own compiled Extensions/atomic_distances.pyx
(oracle/_ref, built by oracle/build_ref.py from /root/reference).  Run in the build container:

    python tests/gen_golden_atomic_distances.py

The fixture holds small systems in every geometry (orthorhombic, triclinic, unwrapped, non-periodic), their type-pair
limits and the outputs (nintra, dintra, ninter, dinter) of multiple_atomic_distances_coords / full_atomic_distances_coords
for every flag combination the distance constraints use (Constraints/DistanceConstraints.py)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402

FLAG_SETS = [dict(), dict(countWithinLimits=False), dict(reduceDistance=True), dict(reduceDistanceToUpper=True),
             dict(reduceDistanceToLower=True, interMolecular=False), dict(intraMolecular=False),
             dict(countWithinLimits=False, reduceDistance=True, intraMolecular=False)]


def systems():
    rng = np.random.default_rng(2611)
    tri = np.array([[19, 0, 0], [2.5, 18, 0], [-1.5, 3, 17]], np.float32)
    out = []
    for name, n, nT, basis, pbc, spread, molsize in (("ortho", 900, 3, np.diag([18.0, 19.0, 17.0]).astype(np.float32), True, 0.0, 1),
                                                     ("tri_molecular", 780, 2, tri, True, 0.0, 13),
                                                     ("tri_unwrapped", 600, 4, tri, True, 1.3, 5),
                                                     ("non_periodic", 700, 2, np.eye(3, dtype=np.float32), False, 0.0, 4)):
        box = (rng.random((n, 3)) * (1 + 2 * spread) - spread).astype(np.float32)
        if not pbc:
            box = (box * 17.0).astype(np.float32)
        el = rng.integers(0, nT, n).astype(np.int32)
        mol = (np.arange(n) // molsize).astype(np.int32)
        lo = (rng.random((nT, nT, 1)) * 1.2).astype(np.float32)
        up = (lo + 0.8 + rng.random((nT, nT, 1)) * 2.5).astype(np.float32)
        idx = rng.integers(0, n, 11).astype(np.int32)
        out.append((name, box, basis, pbc, mol, el, nT, lo, up, idx))
    return out


def main():
    assert build_ref.build(), "cannot build oracle/_ref"
    build_ref.load()
    ad = importlib.import_module("fullrmc.Core.atomic_distances")
    out = {"names": np.array([s[0] for s in systems()]), "n_flag_sets": np.int32(len(FLAG_SETS))}
    for name, box, basis, pbc, mol, el, nT, lo, up, idx in systems():
        out.update({name + "/boxCoords": box, name + "/basis": basis, name + "/isPBC": np.bool_(pbc), name + "/moleculeIndex": mol,
                    name + "/elementIndex": el, name + "/numberOfElements": np.int32(nT), name + "/lowerLimit": lo,
                    name + "/upperLimit": up, name + "/indexes": idx})
        common = dict(boxCoords=box, basis=basis, isPBC=pbc, moleculeIndex=mol, elementIndex=el, numberOfElements=nT,
                      lowerLimit=lo, upperLimit=up, ncores=1)
        for fi, flags in enumerate(FLAG_SETS):
            for allAtoms in (True, False):
                r = ad.multiple_atomic_distances_coords(indexes=idx, allAtoms=allAtoms, **common, **flags)
                for key, a in zip(("nintra", "dintra", "ninter", "dinter"), r):
                    out["%s/multiple/%d/%d/%s" % (name, fi, int(allAtoms), key)] = np.asarray(a).copy()
            r = ad.full_atomic_distances_coords(**common, **flags)
            for key, a in zip(("nintra", "dintra", "ninter", "dinter"), r):
                out["%s/full/%d/%s" % (name, fi, key)] = np.asarray(a).copy()
    path = os.path.join(ROOT, "tests", "golden", "atomic_distances.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
