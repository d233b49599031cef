"""ctypes front-end of oracle/pairhist_oracle.c -- TEST INFRASTRUCTURE ONLY.

Function names, argument names and return conventions mirror the reference's
extension modules (pairs_histograms.pyx:77,150,225,289,343; pairs_distances.pyx:827,874;
reciprocal_space.pyx:42,82) so parity tests read like calls into the reference.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpairhist_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force=False):
    src = os.path.join(_HERE, "pairhist_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "CC=/usr/bin/gcc", "libpairhist_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t)


def set_emulate_spill(on):
    """reproduce (True) or drop (False, default) the reference's out-of-bounds write when bin == histSize"""
    lib().orc_set_emulate_spill(int(bool(on)))


def max_threads():
    return int(lib().orc_max_threads())


def pairs_distances_to_indexcoords(atomIndex, coords, basis, isPBC, allAtoms=True, ncores=1):
    coords, basis = _f32(coords), _f32(basis)
    n = coords.shape[0]
    out = np.zeros(n, dtype=np.float32)
    lib().orc_pairs_distances_to_indexcoords(ctypes.c_int32(int(atomIndex)), _p(coords, _f32p), ctypes.c_int64(n),
                                             _p(basis, _f32p), int(bool(isPBC)), int(bool(allAtoms)), _p(out, _f32p))
    return out


def pairs_distances_to_point(point, coords, basis, isPBC, ncores=1):
    point, coords, basis = _f32(point), _f32(coords), _f32(basis)
    n = coords.shape[0]
    out = np.zeros(n, dtype=np.float32)
    lib().orc_pairs_distances_to_point(_p(point, _f32p), _p(coords, _f32p), ctypes.c_int64(n), _p(basis, _f32p),
                                       int(bool(isPBC)), _p(out, _f32p))
    return out


def pairs_differences_to_point(point, coords, basis, isPBC, ncores=1):
    point, coords, basis = _f32(point), _f32(coords), _f32(basis)
    n = coords.shape[0]
    out = np.zeros((n, 3), dtype=np.float32)
    lib().orc_pairs_differences(_p(point, _f32p), _p(coords, _f32p), ctypes.c_int64(n), _p(basis, _f32p),
                                int(bool(isPBC)), 1, ctypes.c_int64(0), _p(out, _f32p))
    return out


def pairs_differences_to_indexcoords(atomIndex, coords, basis, isPBC, allAtoms=True, ncores=1):
    coords, basis = _f32(coords), _f32(basis)
    n = coords.shape[0]
    out = np.zeros((n, 3), dtype=np.float32)
    point = coords[int(atomIndex)].copy()
    start = 0 if allAtoms else int(atomIndex)
    lib().orc_pairs_differences(_p(point, _f32p), _p(coords, _f32p), ctypes.c_int64(n), _p(basis, _f32p),
                                int(bool(isPBC)), 1, ctypes.c_int64(start), _p(out, _f32p))
    return out


def single_pairs_histograms(atomIndex, distances, moleculeIndex, elementIndex, hintra, hinter,
                            minDistance, maxDistance, bin, allAtoms=True, ncores=1):
    assert hintra.dtype == np.float32 and hintra.flags.c_contiguous
    assert hinter.dtype == np.float32 and hinter.flags.c_contiguous
    distances = _f32(distances)
    mol, el = _i32(moleculeIndex), _i32(elementIndex)
    ov = ctypes.c_uint64(0)
    lib().orc_single_pairs_histograms(ctypes.c_int32(int(atomIndex)), _p(distances, _f32p), ctypes.c_int64(1),
                                      ctypes.c_int64(distances.shape[0]), _p(mol, _i32p), _p(el, _i32p),
                                      int(hintra.shape[0]), int(hintra.shape[2]), _p(hintra, _f32p), _p(hinter, _f32p),
                                      ctypes.c_float(minDistance), ctypes.c_float(maxDistance), ctypes.c_float(bin),
                                      int(bool(allAtoms)), ctypes.byref(ov))
    return int(ov.value)


def multiple_pairs_histograms_coords(indexes, boxCoords, basis, isPBC, moleculeIndex, elementIndex,
                                     numberOfElements, minDistance, maxDistance, bin, histSize,
                                     allAtoms=True, ncores=1, return_overflow=False):
    indexes, coords, basis = _i32(indexes), _f32(boxCoords), _f32(basis)
    mol, el = _i32(moleculeIndex), _i32(elementIndex)
    nEl, hs = int(numberOfElements), int(histSize)
    hintra = np.zeros((nEl, nEl, hs), dtype=np.float32)
    hinter = np.zeros((nEl, nEl, hs), dtype=np.float32)
    ov = ctypes.c_uint64(0)
    lib().orc_multiple_pairs_histograms_coords(_p(indexes, _i32p), ctypes.c_int64(indexes.shape[0]), _p(coords, _f32p),
                                               ctypes.c_int64(coords.shape[0]), _p(basis, _f32p), int(bool(isPBC)),
                                               _p(mol, _i32p), _p(el, _i32p), nEl, ctypes.c_float(minDistance),
                                               ctypes.c_float(maxDistance), ctypes.c_float(bin), hs, int(bool(allAtoms)),
                                               _p(hintra, _f32p), _p(hinter, _f32p), ctypes.byref(ov), int(ncores))
    if return_overflow:
        return hintra, hinter, int(ov.value)
    return hintra, hinter


def full_pairs_histograms_coords(boxCoords, basis, isPBC, moleculeIndex, elementIndex, numberOfElements,
                                 minDistance, maxDistance, bin, histSize, ncores=1, return_overflow=False):
    coords = _f32(boxCoords)
    return multiple_pairs_histograms_coords(np.arange(coords.shape[0], dtype=np.int32), coords, basis, isPBC,
                                            moleculeIndex, elementIndex, numberOfElements, minDistance,
                                            maxDistance, bin, histSize, allAtoms=False, ncores=ncores,
                                            return_overflow=return_overflow)


def multiple_pairs_histograms_dists(indexes, distances, moleculeIndex, elementIndex, numberOfElements,
                                    minDistance, maxDistance, bin, histSize, allAtoms=True, ncores=1):
    indexes, distances = _i32(indexes), _f32(distances)
    mol, el = _i32(moleculeIndex), _i32(elementIndex)
    nEl, hs = int(numberOfElements), int(histSize)
    hintra = np.zeros((nEl, nEl, hs), dtype=np.float32)
    hinter = np.zeros((nEl, nEl, hs), dtype=np.float32)
    ov = ctypes.c_uint64(0)
    lib().orc_multiple_pairs_histograms_dists(_p(indexes, _i32p), ctypes.c_int64(indexes.shape[0]), _p(distances, _f32p),
                                              ctypes.c_int64(distances.shape[0]), _p(mol, _i32p), _p(el, _i32p), nEl,
                                              ctypes.c_float(minDistance), ctypes.c_float(maxDistance),
                                              ctypes.c_float(bin), hs, int(bool(allAtoms)),
                                              _p(hintra, _f32p), _p(hinter, _f32p), ctypes.byref(ov))
    return hintra, hinter


def full_pairs_histograms_dists(distances, moleculeIndex, elementIndex, numberOfElements,
                                minDistance, maxDistance, bin, histSize, ncores=1):
    distances = _f32(distances)
    return multiple_pairs_histograms_dists(np.arange(distances.shape[1], dtype=np.int32), distances, moleculeIndex,
                                           elementIndex, numberOfElements, minDistance, maxDistance, bin, histSize,
                                           allAtoms=False)


def Gr_to_sq(distances, Gr, qrange):
    distances, Gr, qrange = _f32(distances), _f32(Gr), _f32(qrange)
    sq = np.zeros(qrange.shape[0], dtype=np.float32)
    lib().orc_Gr_to_sq(_p(distances, _f32p), _p(Gr, _f32p), ctypes.c_int64(distances.shape[0]), _p(qrange, _f32p),
                       ctypes.c_int64(qrange.shape[0]), _p(sq, _f32p))
    return sq


def gr_to_sq(distances, gr, qrange, rho):
    distances, gr, qrange = _f32(distances), _f32(gr), _f32(qrange)
    sq = np.zeros(qrange.shape[0], dtype=np.float32)
    lib().orc_gr_to_sq(_p(distances, _f32p), _p(gr, _f32p), ctypes.c_int64(distances.shape[0]), _p(qrange, _f32p),
                       ctypes.c_int64(qrange.shape[0]), ctypes.c_float(rho), _p(sq, _f32p))
    return sq


# ---------------------------------------------------------------- atomic distances (Extensions/atomic_distances.pyx)
def _flags(interMolecular, intraMolecular, countWithinLimits, reduceDistanceToUpper, reduceDistanceToLower, reduceDistance):
    return (int(bool(interMolecular)) | int(bool(intraMolecular)) << 1 | int(bool(countWithinLimits)) << 2 |
            int(bool(reduceDistanceToUpper)) << 3 | int(bool(reduceDistanceToLower)) << 4 | int(bool(reduceDistance)) << 5)


def multiple_atomic_distances_coords(indexes, boxCoords, basis, isPBC, moleculeIndex, elementIndex, numberOfElements,
                                     lowerLimit, upperLimit, interMolecular=True, intraMolecular=True, countWithinLimits=True,
                                     reduceDistanceToUpper=False, reduceDistanceToLower=False, reduceDistance=False,
                                     allAtoms=True, ncores=1):
    """atomic_distances.pyx:326-417; returns (nintra, dintra, ninter, dinter), each [nT, nT, 1]"""
    indexes, coords, basis = _i32(indexes), _f32(boxCoords), _f32(basis)
    mol, el = _i32(moleculeIndex), _i32(elementIndex)
    nT = int(numberOfElements)
    lo, up = _f32(lowerLimit).reshape(-1), _f32(upperLimit).reshape(-1)
    assert lo.shape[0] == nT * nT and up.shape[0] == nT * nT
    nintra = np.zeros((nT, nT, 1), np.int32); ninter = np.zeros((nT, nT, 1), np.int32)
    dintra = np.zeros((nT, nT, 1), np.float32); dinter = np.zeros((nT, nT, 1), np.float32)
    fn = lib().orc_multiple_atomic_distances_coords
    fn.restype = None
    fn(_p(indexes, _i32p), ctypes.c_int64(indexes.shape[0]), _p(coords, _f32p), ctypes.c_int64(coords.shape[0]), _p(basis, _f32p),
       int(bool(isPBC)), _p(mol, _i32p), _p(el, _i32p), nT, _p(lo, _f32p), _p(up, _f32p),
       _flags(interMolecular, intraMolecular, countWithinLimits, reduceDistanceToUpper, reduceDistanceToLower, reduceDistance),
       int(bool(allAtoms)), _p(nintra, _i32p), _p(dintra, _f32p), _p(ninter, _i32p), _p(dinter, _f32p))
    return nintra, dintra, ninter, dinter


def full_atomic_distances_coords(boxCoords, basis, isPBC, moleculeIndex, elementIndex, numberOfElements, lowerLimit, upperLimit,
                                 interMolecular=True, intraMolecular=True, reduceDistanceToUpper=False, reduceDistanceToLower=False,
                                 reduceDistance=False, countWithinLimits=True, ncores=1):
    """atomic_distances.pyx:500-567"""
    coords = _f32(boxCoords)
    return multiple_atomic_distances_coords(np.arange(coords.shape[0], dtype=np.int32), coords, basis, isPBC, moleculeIndex,
                                            elementIndex, numberOfElements, lowerLimit, upperLimit, interMolecular=interMolecular,
                                            intraMolecular=intraMolecular, countWithinLimits=countWithinLimits,
                                            reduceDistanceToUpper=reduceDistanceToUpper, reduceDistanceToLower=reduceDistanceToLower,
                                            reduceDistance=reduceDistance, allAtoms=False, ncores=ncores)
