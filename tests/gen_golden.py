"""Generates tests/golden/stateless.npz from the REAL reference kernels (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/gen_golden.py
The outputs are what fullrmc's own compiled Cython functions return on the seeded inputs of
tests/cases.py; the inputs are stored alongside so the fixture is self-contained.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import build_ref  # noqa: E402
import cases as C  # noqa: E402


def main():
    assert build_ref.build(), "cannot build oracle/_ref (is /root/reference mounted?)"
    pd, ph, rs = build_ref.load()
    out = {}
    rng = np.random.default_rng(7)
    for case in C.make_cases():
        nm = case["name"]
        kw = {k: case[k] for k in ("basis", "isPBC", "moleculeIndex", "elementIndex", "numberOfElements",
                                   "minDistance", "maxDistance", "bin", "histSize")}
        for k, v in case.items():
            if k != "name":
                out["%s/in/%s" % (nm, k)] = np.asarray(v)
        hi, he = ph.full_pairs_histograms_coords(boxCoords=case["boxCoords"], **kw)
        out[nm + "/full/intra"], out[nm + "/full/inter"] = hi, he
        n = case["boxCoords"].shape[0]
        idx = C.group_for(case, rng)
        out[nm + "/multi/indexes"] = idx
        for allAtoms in (True, False):
            hi, he = ph.multiple_pairs_histograms_coords(indexes=idx, boxCoords=case["boxCoords"], allAtoms=allAtoms, **kw)
            out["%s/multi/%d/intra" % (nm, allAtoms)], out["%s/multi/%d/inter" % (nm, allAtoms)] = hi, he
        a = int(idx[0])
        out[nm + "/dist/atom"] = np.int32(a)
        out[nm + "/dist/all"] = pd.pairs_distances_to_indexcoords(atomIndex=a, coords=case["boxCoords"],
                                                                  basis=case["basis"], isPBC=case["isPBC"])
        out[nm + "/diff/all"] = pd.pairs_differences_to_indexcoords(atomIndex=a, coords=case["boxCoords"],
                                                                    basis=case["basis"], isPBC=case["isPBC"])
        point = (case["boxCoords"][a] + np.float32(0.37)).astype(np.float32)
        out[nm + "/point"] = point
        out[nm + "/dist/point"] = pd.pairs_distances_to_point(point=point, coords=case["boxCoords"],
                                                              basis=case["basis"], isPBC=case["isPBC"])
        out[nm + "/diff/point"] = pd.pairs_differences_to_point(point=point, coords=case["boxCoords"],
                                                                basis=case["basis"], isPBC=case["isPBC"])
    # reciprocal space (reciprocal_space.pyx:42-109)
    r = np.arange(0.01, 12.0, 0.03, dtype=np.float32)
    G = (np.sin(3.1 * r) * np.exp(-0.2 * r)).astype(np.float32)
    q = np.linspace(0.4, 18.0, 120).astype(np.float32)
    out["recip/r"], out["recip/G"], out["recip/q"] = r, G, q
    out["recip/Gr_to_sq"] = rs.Gr_to_sq(r, G, q)
    out["recip/gr_to_sq"] = rs.gr_to_sq(r, (G * np.float32(0.1) + np.float32(1)).astype(np.float32), q, np.float32(0.085))
    path = os.path.join(ROOT, "tests", "golden", "stateless.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
