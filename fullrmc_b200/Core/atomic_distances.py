"""Drop-in for the coordinate entry points of ``fullrmc.Core.atomic_distances`` (reference:
Extensions/atomic_distances.pyx), the hot loop of InterMolecularDistanceConstraint and its relatives
(Constraints/DistanceConstraints.py:552-737).

``multiple_atomic_distances_coords`` (:326-417) and ``full_atomic_distances_coords`` (:500-567) keep the
reference's names, keyword names, dtypes and return convention ``(nintra, dintra, ninter, dinter)``, each a new
``[numberOfElements, numberOfElements, 1]`` array (int32 counts, float32 distance sums).  Results are bit-identical
to the reference, float sums included: the device adds the hits in the reference's loop order
(csrc/atomdist.cu).  ``ncores`` is accepted and ignored.
"""
import numpy as np

from .. import _lib as L

_F32, _I32 = np.float32, np.int32


def _flags(interMolecular, intraMolecular, countWithinLimits, reduceDistanceToUpper, reduceDistanceToLower, reduceDistance):
    return (int(bool(interMolecular)) | int(bool(intraMolecular)) << 1 | int(bool(countWithinLimits)) << 2 |
            int(bool(reduceDistanceToUpper)) << 3 | int(bool(reduceDistanceToLower)) << 4 | int(bool(reduceDistance)) << 5)


def _limits(lowerLimit, upperLimit, nT):
    lo = L.as_array(lowerLimit, "lowerLimit", _F32, 3)
    up = L.as_array(upperLimit, "upperLimit", _F32, 3)
    for name, a in (("lowerLimit", lo), ("upperLimit", up)):
        # the reference's own assertions (atomic_distances.pyx:369-371)
        assert a.shape[0] == a.shape[1] and a.shape[0] == nT, \
            "%s array must have numberOfElements columns and numberOfElements rows" % name
        assert a.shape[2] == 1, "%s array third dimension must have a length of exactly 1" % name
    return np.ascontiguousarray(lo), np.ascontiguousarray(up)


def multiple_atomic_distances_coords(indexes, boxCoords, basis, isPBC, moleculeIndex, elementIndex, numberOfElements,
                                     lowerLimit, upperLimit, interMolecular=True, intraMolecular=True, countWithinLimits=True,
                                     reduceDistanceToUpper=False, reduceDistanceToLower=False, reduceDistance=False,
                                     allAtoms=True, ncores=1):
    """atomic_distances.pyx:326-417 (same positional order of the optional flags as the reference, :335-342)"""
    lib = L.load_library()
    idx = L.as_array(indexes, "indexes", _I32, 1)
    coords = L.as_array(boxCoords, "boxCoords", _F32, 2)
    b = L.as_array(basis, "basis", _F32, 2)
    mol = L.as_array(moleculeIndex, "moleculeIndex", _I32, 1)
    el = L.as_array(elementIndex, "elementIndex", _I32, 1)
    nT = int(numberOfElements)
    lo, up = _limits(lowerLimit, upperLimit, nT)
    n = coords.shape[0]
    if mol.shape[0] != n or el.shape[0] != n:
        raise ValueError("moleculeIndex/elementIndex length must equal the number of atoms (%d)" % n)
    nintra = np.zeros((nT, nT, 1), _I32); ninter = np.zeros((nT, nT, 1), _I32)
    dintra = np.zeros((nT, nT, 1), _F32); dinter = np.zeros((nT, nT, 1), _F32)
    rc = lib.frmc_multiple_atomic_distances_coords(
        L.device_index(), L.ptr(idx, L.c_i32p), idx.shape[0], L.ptr(coords, L.c_f32p), n, L.ptr(b, L.c_f32p), int(bool(isPBC)),
        L.ptr(mol, L.c_i32p), L.ptr(el, L.c_i32p), nT, L.ptr(lo, L.c_f32p), L.ptr(up, L.c_f32p),
        _flags(interMolecular, intraMolecular, countWithinLimits, reduceDistanceToUpper, reduceDistanceToLower, reduceDistance),
        int(bool(allAtoms)), L.ptr(nintra, L.c_i32p), L.ptr(dintra, L.c_f32p), L.ptr(ninter, L.c_i32p), L.ptr(dinter, L.c_f32p))
    L.check(rc, "multiple_atomic_distances_coords")
    return nintra, dintra, ninter, dinter


def full_atomic_distances_coords(boxCoords, basis, isPBC, moleculeIndex, elementIndex, numberOfElements, lowerLimit, upperLimit,
                                 interMolecular=True, intraMolecular=True, reduceDistanceToUpper=False, reduceDistanceToLower=False,
                                 reduceDistance=False, countWithinLimits=True, ncores=1):
    """atomic_distances.pyx:500-567 (optional flags in the reference's positional order, :508-514) -- in within-limits mode the device sweeps only the block pairs of the k-d ordered
    store that can reach the largest upper limit (csrc/atomdist.cu: full_culled)"""
    lib = L.load_library()
    coords = L.as_array(boxCoords, "boxCoords", _F32, 2)
    b = L.as_array(basis, "basis", _F32, 2)
    mol = L.as_array(moleculeIndex, "moleculeIndex", _I32, 1)
    el = L.as_array(elementIndex, "elementIndex", _I32, 1)
    nT = int(numberOfElements)
    lo, up = _limits(lowerLimit, upperLimit, nT)
    n = coords.shape[0]
    if mol.shape[0] != n or el.shape[0] != n:
        raise ValueError("moleculeIndex/elementIndex length must equal the number of atoms (%d)" % n)
    nintra = np.zeros((nT, nT, 1), _I32); ninter = np.zeros((nT, nT, 1), _I32)
    dintra = np.zeros((nT, nT, 1), _F32); dinter = np.zeros((nT, nT, 1), _F32)
    rc = lib.frmc_full_atomic_distances_coords(
        L.device_index(), L.ptr(coords, L.c_f32p), n, L.ptr(b, L.c_f32p), int(bool(isPBC)), L.ptr(mol, L.c_i32p), L.ptr(el, L.c_i32p),
        nT, L.ptr(lo, L.c_f32p), L.ptr(up, L.c_f32p),
        _flags(interMolecular, intraMolecular, countWithinLimits, reduceDistanceToUpper, reduceDistanceToLower, reduceDistance),
        L.ptr(nintra, L.c_i32p), L.ptr(dintra, L.c_f32p), L.ptr(ninter, L.c_i32p), L.ptr(dinter, L.c_f32p))
    L.check(rc, "full_atomic_distances_coords")
    return nintra, dintra, ninter, dinter


def pair_elements_stats(elementIndex, moleculeIndex, numberOfElements):
    """atomic_distances.pyx:632-672 -- number of ordered pairs (i < j) per (element of i, element of j), split into
    same-molecule and different-molecule pairs; no distances are involved.  The reference walks all N(N-1)/2 pairs;
    the counts only depend on how many atoms of each element precede an atom (overall, and inside its molecule), so
    they are formed here from prefix counts in O(N * numberOfElements) on the host -- set-up work of
    ``set_type_definition`` (DistanceConstraints.py:525-541), not part of a Monte-Carlo step.  Accumulated in int64
    and returned as int32 like the reference's arrays (the same wrap-around for systems beyond 2^31 pairs of a kind)."""
    el = L.as_array(elementIndex, "elementIndex", _I32, 1)
    mol = L.as_array(moleculeIndex, "moleculeIndex", _I32, 1)
    nT = int(numberOfElements)
    n = el.shape[0]
    if mol.shape[0] != n:
        raise ValueError("moleculeIndex length must equal the number of atoms (%d)" % n)
    total = np.zeros((nT, nT), np.int64)
    intra = np.zeros((nT, nT), np.int64)
    if n > 1:
        onehot = (el[:, None] == np.arange(nT, dtype=_I32)[None, :])            # [n, nT]
        before = np.cumsum(onehot, axis=0, dtype=np.int64) - onehot             # atoms of each element strictly before j
        total = before.T @ onehot.astype(np.int64)                              # [a, b] = sum_j [el_j = b] * #(i < j, el_i = a)
        order = np.argsort(mol, kind="stable")                                  # molecule by molecule, original order inside
        oh = onehot[order]
        cs = np.cumsum(oh, axis=0, dtype=np.int64) - oh
        first = np.r_[True, mol[order][1:] != mol[order][:-1]]                  # first atom of each molecule in the sorted list
        base = cs[first][np.cumsum(first) - 1]                                  # prefix counts at the start of the atom's molecule
        intra = (cs - base).T @ oh.astype(np.int64)
    inter = total - intra
    return (np.ascontiguousarray(intra.astype(_I32).reshape(nT, nT, 1)),
            np.ascontiguousarray(inter.astype(_I32).reshape(nT, nT, 1)))
