"""Randomised soak of the rigid pre-filters on the device store (csrc/storecoord.cu, csrc/storedist.cu) against the stateless
kernels (which are pinned to the compiled reference): random systems, definitions / windows, group moves, every other move
applied.  usage: python tools/soak_store_prefilters.py [n_cases] [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic
from fullrmc_b200.Core import atomic_coordination as ac, atomic_distances as ad
from fullrmc_b200.constraints_coordination import _membership
from fullrmc_b200.store import DeviceStore

F32 = np.float32
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
bad = 0
for case in range(n_cases):
    n = int(rng.integers(800, 9000))
    nEl = int(rng.integers(1, 5))
    pbc = bool(rng.random() < 0.7)
    edge = float(rng.uniform(18, 45))
    basis = (np.diag(edge * rng.uniform(0.8, 1.2, 3)) + np.tril(rng.uniform(-0.15, 0.15, (3, 3)) * edge, -1)).astype(F32)
    s = synthetic.random_system(n, int(rng.integers(1, 10**6)), basis, n_elements=nEl, molecule_size=int(rng.choice([1, 2, 4])), isPBC=pbc,
                                spread=(0.3 if (pbc and rng.random() < 0.4) else None))
    ndef = int(rng.integers(1, 7))
    cores, shells, lower, upper = [], [], [], []
    for d in range(ndef):
        cores.append(np.sort(rng.choice(n, int(rng.integers(1, n)), replace=False)).astype(np.int32))
        shells.append(np.sort(rng.choice(n, int(rng.integers(1, n)), replace=False)).astype(np.int32))
        lo = float(rng.choice([0.0, 0.0, 0.8, 1.5]))
        lower.append(F32(lo)); upper.append(F32(lo + rng.uniform(0.5, 3.5)))
    asc, ins = _membership(cores, n), _membership(shells, n)
    nT = nEl
    lo_w = np.zeros((nT, nT, 1), F32)
    up_w = (1.2 + rng.random((nT, nT, 1))).astype(F32); up_w = ((up_w + up_w.transpose(1, 0, 2)) / 2).astype(F32)
    flags = dict(interMolecular=True, intraMolecular=bool(rng.random() < 0.5), reduceDistance=False, reduceDistanceToUpper=True,
                 reduceDistanceToLower=False, countWithinLimits=True)
    ok = True
    with DeviceStore(s.boxCoords, s.basis, pbc, s.moleculeIndex, s.elementIndex, nEl) as st:
        kid = st.coordination_add(cores, shells, lower, upper)
        did = st.distance_add(s.elementIndex, nT, lo_w, up_w, **flags)
        box = s.boxCoords.copy()
        dkw = dict(basis=s.basis, isPBC=pbc, numberOfElements=nT, lowerLimit=lo_w, upperLimit=up_w, **flags)
        for step in range(10):
            k = int(rng.choice([1, 1, 3, 8]))
            idx = np.sort(rng.choice(n, k, replace=False)).astype(np.int32)
            scale = 0.03 if pbc else 0.03 * edge
            moved = (box[idx] + rng.normal(0, scale / (edge if pbc else 1.0) * (1.0 if pbc else 1.0), (k, 3))).astype(F32)
            cn = st.coordination_move(kid, idx, moved).copy()
            counts, sums = st.distance_move(did, idx, moved)
            counts, sums = counts.copy(), sums.copy()
            after = box.copy(); after[idx] = moved
            for which, coords in ((0, box), (1, after)):
                want = np.zeros(ndef, F32)
                ac.multi_atoms_coord_number_coords(indexes=idx, boxCoords=coords, basis=s.basis, isPBC=pbc, coresIndexes=cores, shellsIndexes=shells,
                                                   lowerShells=lower, upperShells=upper, asCoreDefIdxs=asc, inShellDefIdxs=ins, coordNumData=want)
                ok = ok and np.array_equal(cn[which].astype(F32), want)
                ni, di, ne, de = ad.multiple_atomic_distances_coords(indexes=idx, boxCoords=coords, moleculeIndex=s.moleculeIndex,
                                                                     elementIndex=s.elementIndex, allAtoms=True, **dkw)
                ok = ok and np.array_equal(counts[2 * which, 0], ni) and np.array_equal(counts[2 * which, 1], ne)
                ok = ok and np.array_equal(sums[2 * which, 0], di) and np.array_equal(sums[2 * which, 1], de)
            if step % 2 == 0:
                st.move_atoms(idx, moved); box = after
    bad += 0 if ok else 1
    print("case %2d n %5d nEl %d pbc %-5s defs %d  %s" % (case, n, nEl, pbc, ndef, "ok" if ok else "MISMATCH"), flush=True)
print("soak: %d cases, %d mismatches" % (n_cases, bad))
sys.exit(1 if bad else 0)
