"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``fullrmc.Core.atomic_coordination`` (reference:
Extensions/atomic_coordination.pyx).  Never imported by the product
(fullrmc_b200/); only tests/ may use it.

The Python-level structure (which definition is visited for which atom, in which order, and what is added to
``coordNumData``) follows the reference function by function; the two inner loops are in oracle/pairhist_oracle.c
(orc_single_atom_single_shell_coords / _dists).  Pinned against the reference's own compiled module by
tests/golden/atomic_coordination.npz (tests/gen_golden_atomic_coordination.py).
"""
import ctypes

import numpy as np

from . import pairhist as _ph

_F32, _I32 = np.float32, np.int32
_f32p, _i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
_bound = False


def _lib():
    global _bound
    lib = _ph.lib()
    if not _bound:
        lib.orc_single_atom_single_shell_coords.restype = ctypes.c_float
        lib.orc_single_atom_single_shell_coords.argtypes = [ctypes.c_int32, _i32p, ctypes.c_int64, _f32p, ctypes.c_int64, _f32p,
                                                            ctypes.c_int, ctypes.c_float, ctypes.c_float]
        lib.orc_single_atom_single_shell_dists.restype = ctypes.c_float
        lib.orc_single_atom_single_shell_dists.argtypes = [_f32p, ctypes.c_int64, _i32p, ctypes.c_int64, ctypes.c_float, ctypes.c_float]
        _bound = True
    return lib


def _p(a, t):
    return a.ctypes.data_as(t)


def single_atom_single_shell_subdists(distances, lowerShell, upperShell, ncores=1):
    """atomic_coordination.pyx:55-67"""
    d = np.ascontiguousarray(distances, _F32)
    return float(_lib().orc_single_atom_single_shell_dists(_p(d, _f32p), d.shape[0], None, d.shape[0], lowerShell, upperShell))


def single_atom_single_shell_totdists(distances, shellIndexes, lowerShell, upperShell, ncores=1):
    """atomic_coordination.pyx:71-85"""
    d = np.ascontiguousarray(distances, _F32)
    s = np.ascontiguousarray(shellIndexes, _I32)
    return float(_lib().orc_single_atom_single_shell_dists(_p(d, _f32p), d.shape[0], _p(s, _i32p), s.shape[0], lowerShell, upperShell))


def single_atom_single_shell_coords(coreIndex, shellIndexes, boxCoords, basis, isPBC, lowerShell, upperShell, ncores=1):
    """atomic_coordination.pyx:89-112"""
    c = np.ascontiguousarray(boxCoords, _F32)
    b = np.ascontiguousarray(basis, _F32)
    s = np.ascontiguousarray(shellIndexes, _I32)
    return float(_lib().orc_single_atom_single_shell_coords(int(coreIndex), _p(s, _i32p), s.shape[0], _p(c, _f32p), c.shape[0],
                                                            _p(b, _f32p), int(bool(isPBC)), lowerShell, upperShell))


def single_atom_multi_shells_totdists(distances, shellsIndexes, lowerShells, upperShells, ncores=1):
    """atomic_coordination.pyx:116-135"""
    out = np.zeros(len(shellsIndexes), _F32)
    for i in range(len(shellsIndexes)):
        out[i] = single_atom_single_shell_totdists(distances, shellsIndexes[i], lowerShells[i], upperShells[i])
    return out


def single_atom_multi_shells_coords(coreIndex, shellsIndexes, boxCoords, basis, isPBC, lowerShells, upperShells, ncores=1):
    """atomic_coordination.pyx:139-167"""
    out = np.zeros(len(shellsIndexes), _F32)
    for i in range(len(shellsIndexes)):
        out[i] = single_atom_single_shell_coords(coreIndex, shellsIndexes[i], boxCoords, basis, isPBC, lowerShells[i], upperShells[i])
    return out


def single_atom_coord_number_totdists(atomIndex, distances, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs,
                                      inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:171-198 -- both loops walk asCoreDefIdxs, as the reference does (:185, :192)"""
    for defIdx in asCoreDefIdxs[atomIndex]:
        coordNumData[defIdx] += _F32(single_atom_single_shell_totdists(distances, shellsIndexes[defIdx], lowerShells[defIdx], upperShells[defIdx]))
    for defIdx in asCoreDefIdxs[atomIndex]:
        coordNumData[defIdx] += _F32(single_atom_single_shell_totdists(distances, coresIndexes[defIdx], lowerShells[defIdx], upperShells[defIdx]))


def single_atom_coord_number_coords(atomIndex, boxCoords, basis, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells,
                                    asCoreDefIdxs, inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:207-240"""
    for defIdx in asCoreDefIdxs[atomIndex]:
        coordNumData[defIdx] += _F32(single_atom_single_shell_coords(atomIndex, shellsIndexes[defIdx], boxCoords, basis, isPBC,
                                                                     lowerShells[defIdx], upperShells[defIdx]))
    for defIdx in inShellDefIdxs[atomIndex]:
        coordNumData[defIdx] += _F32(single_atom_single_shell_coords(atomIndex, coresIndexes[defIdx], boxCoords, basis, isPBC,
                                                                     lowerShells[defIdx], upperShells[defIdx]))


def multi_atoms_coord_number_totdists(indexes, distances, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs,
                                      inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:249-276"""
    for i in range(len(indexes)):
        single_atom_coord_number_totdists(int(indexes[i]), distances[i], coresIndexes, shellsIndexes, lowerShells, upperShells,
                                          asCoreDefIdxs, inShellDefIdxs, coordNumData)


def multi_atoms_coord_number_coords(indexes, boxCoords, basis, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells,
                                    asCoreDefIdxs, inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:280-313"""
    for i in range(len(indexes)):
        single_atom_coord_number_coords(int(indexes[i]), boxCoords, basis, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells,
                                        asCoreDefIdxs, inShellDefIdxs, coordNumData)


def all_atoms_coord_number_totdists(distances, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs, inShellDefIdxs,
                                    coordNumData, ncores=1):
    """atomic_coordination.pyx:317-345 -- unusable in the reference: it passes its 2-d ndarray on to
    multi_atoms_coord_number_totdists, whose ``distances`` argument is typed ``list``, so every call ends in this
    TypeError (pinned by tests/golden/atomic_coordination.npz); kept for the name, with the same outcome"""
    raise TypeError("Argument 'distances' has incorrect type (expected list, got numpy.ndarray)")



def all_atoms_coord_number_coords(boxCoords, basis, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs,
                                  inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:349-376"""
    multi_atoms_coord_number_coords(np.arange(boxCoords.shape[0], dtype=_I32), boxCoords, basis, isPBC, coresIndexes, shellsIndexes,
                                    lowerShells, upperShells, asCoreDefIdxs, inShellDefIdxs, coordNumData)
