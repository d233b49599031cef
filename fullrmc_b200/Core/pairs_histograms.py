"""Drop-in for ``fullrmc.Core.pairs_histograms`` (reference: Extensions/pairs_histograms.pyx).

Every function keeps the reference's name, positional order, keyword names and return
convention; ``ncores`` is accepted and ignored (the GPU is the parallelism).  Arrays must
be float32 / int32 numpy arrays exactly like the Cython typed-buffer arguments
(``None`` -> TypeError, wrong dtype/ndim -> ValueError).  Results are bit-identical to
the reference (tests/test_gpu_stateless.py).

``LAST_EDGE_OVERFLOW`` holds, after each call, the number of in-range pairs whose fp32
bin index rounded up to ``histSize`` (undefined behaviour in the reference, which has
boundscheck(False); dropped and counted here).
"""
import ctypes

import numpy as np

from .. import _lib as L

LAST_EDGE_OVERFLOW = 0

_F32, _I32 = np.float32, np.int32


def _set_overflow(v):
    global LAST_EDGE_OVERFLOW
    LAST_EDGE_OVERFLOW = int(v)


def _check_system(n, moleculeIndex, elementIndex):
    if moleculeIndex.shape[0] != n or elementIndex.shape[0] != n:
        raise ValueError("moleculeIndex/elementIndex length must equal the number of atoms (%d)" % n)


def single_pairs_histograms(atomIndex, distances, moleculeIndex, elementIndex, hintra, hinter,
                            minDistance, maxDistance, bin, allAtoms=True, ncores=1):
    """pairs_histograms.pyx:77-141 -- updates ``hintra``/``hinter`` IN PLACE from one distance row."""
    lib = L.load_library()
    distances = L.as_array(distances, "distances", _F32, 1)
    mol = L.as_array(moleculeIndex, "moleculeIndex", _I32, 1)
    el = L.as_array(elementIndex, "elementIndex", _I32, 1)
    for name, h in (("hintra", hintra), ("hinter", hinter)):
        if h is None:
            raise TypeError("Argument '%s' must not be None" % name)
        if not isinstance(h, np.ndarray) or h.dtype != _F32 or h.ndim != 3:
            raise ValueError("Buffer dtype mismatch or wrong number of dimensions for '%s'" % name)
    n = distances.shape[0]
    _check_system(n, mol, el)
    hi = np.ascontiguousarray(hintra)
    he = np.ascontiguousarray(hinter)
    nEl, hs = hi.shape[0], hi.shape[2]
    ov = ctypes.c_uint64(0)
    rc = lib.frmc_single_pairs_histograms(L.device_index(), int(atomIndex), L.ptr(distances, L.c_f32p), 1, n,
                                          L.ptr(mol, L.c_i32p), L.ptr(el, L.c_i32p), nEl, hs,
                                          L.ptr(hi, L.c_f32p), L.ptr(he, L.c_f32p),
                                          float(_F32(minDistance)), float(_F32(maxDistance)), float(_F32(bin)),
                                          int(bool(allAtoms)), ctypes.byref(ov))
    L.check(rc, "single_pairs_histograms")
    if hi is not hintra:
        hintra[...] = hi
    if he is not hinter:
        hinter[...] = he
    _set_overflow(ov.value)


def multiple_pairs_histograms_coords(indexes, boxCoords, basis, isPBC, moleculeIndex, elementIndex,
                                     numberOfElements, minDistance, maxDistance, bin, histSize,
                                     allAtoms=True, ncores=1):
    """pairs_histograms.pyx:150-217 -- listed atoms against all (or following) atoms."""
    lib = L.load_library()
    indexes = L.as_array(indexes, "indexes", _I32, 1)
    coords = L.as_array(boxCoords, "boxCoords", _F32, 2)
    basis = L.as_array(basis, "basis", _F32, 2)
    mol = L.as_array(moleculeIndex, "moleculeIndex", _I32, 1)
    el = L.as_array(elementIndex, "elementIndex", _I32, 1)
    n = coords.shape[0]
    if coords.shape[1] != 3 or basis.shape != (3, 3):
        raise ValueError("boxCoords must be (N,3) and basis (3,3)")
    _check_system(n, mol, el)
    nEl, hs = int(numberOfElements), int(histSize)
    hintra = np.empty((nEl, nEl, hs), dtype=_F32)
    hinter = np.empty((nEl, nEl, hs), dtype=_F32)
    ov = ctypes.c_uint64(0)
    rc = lib.frmc_multiple_pairs_histograms_coords(L.device_index(), L.ptr(indexes, L.c_i32p), indexes.shape[0],
                                                   L.ptr(coords, L.c_f32p), n, L.ptr(basis, L.c_f32p), int(bool(isPBC)),
                                                   L.ptr(mol, L.c_i32p), L.ptr(el, L.c_i32p), nEl,
                                                   float(_F32(minDistance)), float(_F32(maxDistance)), float(_F32(bin)),
                                                   hs, int(bool(allAtoms)), L.ptr(hintra, L.c_f32p),
                                                   L.ptr(hinter, L.c_f32p), ctypes.byref(ov))
    L.check(rc, "multiple_pairs_histograms_coords")
    _set_overflow(ov.value)
    return hintra, hinter


def multiple_pairs_histograms_dists(indexes, distances, moleculeIndex, elementIndex, numberOfElements,
                                    minDistance, maxDistance, bin, histSize, allAtoms=True, ncores=1):
    """pairs_histograms.pyx:225-281 -- same binning from an [N,k] distance matrix."""
    lib = L.load_library()
    indexes = L.as_array(indexes, "indexes", _I32, 1)
    distances = L.as_array(distances, "distances", _F32, 2)
    mol = L.as_array(moleculeIndex, "moleculeIndex", _I32, 1)
    el = L.as_array(elementIndex, "elementIndex", _I32, 1)
    n, k = distances.shape
    if k != indexes.shape[0]:
        raise ValueError("distances must have one column per index")
    _check_system(n, mol, el)
    nEl, hs = int(numberOfElements), int(histSize)
    hintra = np.empty((nEl, nEl, hs), dtype=_F32)
    hinter = np.empty((nEl, nEl, hs), dtype=_F32)
    ov = ctypes.c_uint64(0)
    rc = lib.frmc_multiple_pairs_histograms_dists(L.device_index(), L.ptr(indexes, L.c_i32p), k,
                                                  L.ptr(distances, L.c_f32p), n, L.ptr(mol, L.c_i32p),
                                                  L.ptr(el, L.c_i32p), nEl, float(_F32(minDistance)),
                                                  float(_F32(maxDistance)), float(_F32(bin)), hs, int(bool(allAtoms)),
                                                  L.ptr(hintra, L.c_f32p), L.ptr(hinter, L.c_f32p), ctypes.byref(ov))
    L.check(rc, "multiple_pairs_histograms_dists")
    _set_overflow(ov.value)
    return hintra, hinter


def full_pairs_histograms_coords(boxCoords, basis, isPBC, moleculeIndex, elementIndex, numberOfElements,
                                 minDistance, maxDistance, bin, histSize, ncores=1, _shard=0, _nshards=1, _devices=None):
    """pairs_histograms.pyx:289-335 -- ordered upper triangle [el[i], el[j]], i<j, over the
    whole system (CUDA sweep).  Not part of the reference signature: ``_shard/_nshards`` select one slice
    of the row list for one-process-per-GPU runs; ``_devices`` (default: $FULLRMC_B200_DEVICES) spreads ONE call
    over several GPUs of the box, the integer counts combined by an NCCL all-reduce inside the library."""
    lib = L.load_library()
    coords = L.as_array(boxCoords, "boxCoords", _F32, 2)
    basis = L.as_array(basis, "basis", _F32, 2)
    mol = L.as_array(moleculeIndex, "moleculeIndex", _I32, 1)
    el = L.as_array(elementIndex, "elementIndex", _I32, 1)
    n = coords.shape[0]
    if coords.shape[1] != 3 or basis.shape != (3, 3):
        raise ValueError("boxCoords must be (N,3) and basis (3,3)")
    _check_system(n, mol, el)
    nEl, hs = int(numberOfElements), int(histSize)
    hintra = np.empty((nEl, nEl, hs), dtype=_F32)
    hinter = np.empty((nEl, nEl, hs), dtype=_F32)
    ov = ctypes.c_uint64(0)
    devices = L.device_list() if _devices is None else [int(d) for d in _devices]
    if len(devices) > 1 and int(_nshards) == 1:
        devs = (ctypes.c_int * len(devices))(*devices)
        rc = lib.frmc_full_pairs_histograms_coords_multi(len(devices), devs, L.ptr(coords, L.c_f32p), n, L.ptr(basis, L.c_f32p),
                                                         int(bool(isPBC)), L.ptr(mol, L.c_i32p), L.ptr(el, L.c_i32p), nEl,
                                                         float(_F32(minDistance)), float(_F32(maxDistance)), float(_F32(bin)), hs,
                                                         L.ptr(hintra, L.c_f32p), L.ptr(hinter, L.c_f32p), ctypes.byref(ov))
        L.check(rc, "full_pairs_histograms_coords (%d devices)" % len(devices))
        _set_overflow(ov.value)
        return hintra, hinter
    rc = lib.frmc_full_pairs_histograms_coords(devices[0] if _devices is not None else L.device_index(), L.ptr(coords, L.c_f32p), n, L.ptr(basis, L.c_f32p),
                                               int(bool(isPBC)), L.ptr(mol, L.c_i32p), L.ptr(el, L.c_i32p), nEl,
                                               float(_F32(minDistance)), float(_F32(maxDistance)), float(_F32(bin)),
                                               hs, int(_shard), int(_nshards), L.ptr(hintra, L.c_f32p),
                                               L.ptr(hinter, L.c_f32p), ctypes.byref(ov))
    L.check(rc, "full_pairs_histograms_coords")
    _set_overflow(ov.value)
    return hintra, hinter


def full_pairs_histograms_dists(distances, moleculeIndex, elementIndex, numberOfElements,
                                minDistance, maxDistance, bin, histSize, ncores=1):
    """pairs_histograms.pyx:343-383 -- indexes = arange(distances.shape[1]), allAtoms=False."""
    distances = L.as_array(distances, "distances", _F32, 2)
    indexes = np.arange(distances.shape[1], dtype=_I32)
    return multiple_pairs_histograms_dists(indexes=indexes, distances=distances, moleculeIndex=moleculeIndex,
                                           elementIndex=elementIndex, numberOfElements=numberOfElements,
                                           minDistance=minDistance, maxDistance=maxDistance, bin=bin,
                                           histSize=histSize, allAtoms=False, ncores=ncores)
