"""Randomised parity soak of the stateless full histogram (device ordering, unit- and chunk-level culling) against the C
oracle: random geometries (orthorhombic / triclinic / non-periodic, wrapped and unwrapped coordinates, clustered and uniform
densities), random grids.  usage: python tools/soak_fullhist.py [n_cases] [seed]   -> prints a line per case, exits 1 on a mismatch"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200.Core import pairs_histograms as ph
from oracle import pairhist as orc

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
bad = 0
for case in range(n_cases):
    n = int(rng.integers(3000, 45000))
    kind = rng.choice(["ortho", "tri", "ibc"])
    edge = float(rng.uniform(25, 90))
    if kind == "ortho":
        basis = np.diag(edge * rng.uniform(0.8, 1.25, 3)).astype(np.float32)
    else:
        basis = (np.diag(edge * rng.uniform(0.8, 1.25, 3)) + np.tril(rng.uniform(-0.2, 0.2, (3, 3)) * edge, -1)).astype(np.float32)
    pbc = kind != "ibc"
    box = rng.random((n, 3))
    if rng.random() < 0.4:                                   # clustered: a third of the atoms in a small blob
        m = n // 3
        box[:m] = 0.5 + 0.05 * rng.standard_normal((m, 3))
    if pbc and rng.random() < 0.5:
        box = box + rng.integers(-2, 3, (n, 3))              # unwrapped images
    if not pbc:
        box = (box - 0.5) @ basis.astype(np.float64)         # real coordinates around the origin (negative values too)
        basis = np.eye(3, dtype=np.float32)
    box = box.astype(np.float32)
    nEl = int(rng.integers(1, 6))
    el = rng.integers(0, nEl, n).astype(np.int32)
    msize = int(rng.choice([1, 1, 2, 5]))
    mol = (np.arange(n) // msize).astype(np.int32)
    if rng.random() < 0.3:
        perm = rng.permutation(n); box, el, mol = box[perm], el[perm], mol[perm]       # molecules scattered over the index range
    rmin = float(rng.choice([0.0, 0.0, 0.7, 1.3]))
    b = float(rng.choice([0.02, 0.05, 0.1]))
    hs = int(rng.integers(40, 400))
    kw = dict(basis=basis, isPBC=pbc, moleculeIndex=mol, elementIndex=el, numberOfElements=nEl, minDistance=np.float32(rmin),
              maxDistance=np.float32(rmin + b * hs), bin=np.float32(b), histSize=hs)
    t0 = time.perf_counter()
    gi, ge = ph.full_pairs_histograms_coords(boxCoords=box, **kw)
    t1 = time.perf_counter()
    wi, we = orc.full_pairs_histograms_coords(boxCoords=box, ncores=orc.max_threads(), **kw)
    ok = bool(np.array_equal(gi, wi) and np.array_equal(ge, we))
    bad += 0 if ok else 1
    print("case %2d %-5s n %6d nEl %d mol %d rmin %.1f rmax %5.1f hs %3d  hits %.3g  gpu %.1f ms  %s" % (
        case, kind, n, nEl, msize, rmin, rmin + b * hs, hs, float(wi.sum() + we.sum()), 1e3 * (t1 - t0), "ok" if ok else "MISMATCH"), flush=True)
print("soak: %d cases, %d mismatches" % (n_cases, bad))
sys.exit(1 if bad else 0)
