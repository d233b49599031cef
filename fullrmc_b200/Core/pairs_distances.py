"""Drop-in for ``fullrmc.Core.pairs_distances`` (reference: Extensions/pairs_distances.pyx).

All ten public functions of the reference module, same names and keywords.  They share one
CUDA kernel ("k points against N coordinate rows", csrc/stateless.cu) through
``frmc_points_to_coords``.  Rows the reference leaves uninitialised (``allAtoms=False``,
rows below the atom index: np.empty at pairs_distances.pyx:899) are returned as 0.
"""
import ctypes

import numpy as np

from .. import _lib as L

_F32, _I32 = np.float32, np.int32


def _basis(basis):
    if basis is None:
        raise TypeError("Argument 'basis' must not be None")
    b = np.ascontiguousarray(np.asarray(basis), dtype=_F32)   # reference takes a C_FLOAT32[:,:] memoryview
    if b.shape != (3, 3):
        raise ValueError("basis must be (3,3)")
    return b


def _run(points, from_index, start, coords, basis, isPBC, ibc_sign, want_diff):
    lib = L.load_library()
    n = coords.shape[0]
    k = from_index.shape[0] if from_index is not None else points.shape[0]
    out = np.zeros((n, 3, k) if want_diff else (n, k), dtype=_F32)
    st = None if start is None else np.ascontiguousarray(start, dtype=np.int64)
    rc = lib.frmc_points_to_coords(L.device_index(), L.ptr(points, L.c_f32p), L.ptr(from_index, L.c_i32p),
                                   L.ptr(st, L.c_i64p), k, L.ptr(coords, L.c_f32p), n, L.ptr(basis, L.c_f32p),
                                   int(bool(isPBC)), int(ibc_sign), int(want_diff), L.ptr(out, L.c_f32p))
    L.check(rc, "pairs_distances")
    return out


def _coords(a, name="coords"):
    c = L.as_array(a, name, _F32, 2)
    if c.shape[1] != 3:
        raise ValueError("%s must be (N,3)" % name)
    return c


def from_to_points_differences(pointsFrom, pointsTo, basis, isPBC, ncores=1):
    """pairs_distances.pyx:483-525 -- boundaryConditions(pointsTo[i]-pointsFrom[i]) -> (N,3)."""
    lib = L.load_library()
    a, b, basis = _coords(pointsFrom, "pointsFrom"), _coords(pointsTo, "pointsTo"), _basis(basis)
    if a.shape != b.shape:
        raise ValueError("pointsFrom and pointsTo must have the same shape")
    out = np.empty((a.shape[0], 3), dtype=_F32)
    rc = lib.frmc_from_to_points_differences(L.device_index(), L.ptr(a, L.c_f32p), L.ptr(b, L.c_f32p), a.shape[0],
                                             L.ptr(basis, L.c_f32p), int(bool(isPBC)), L.ptr(out, L.c_f32p))
    L.check(rc, "from_to_points_differences")
    return out


def pair_difference_to_point(point1, point2, basis, isPBC, ncores=1):
    """pairs_distances.pyx:534-571 -- boundaryConditions(point2-point1) -> (3,)."""
    p1 = L.as_array(point1, "point1", _F32, 1).reshape(1, 3)
    p2 = L.as_array(point2, "point2", _F32, 1).reshape(1, 3)
    return from_to_points_differences(p1, p2, basis, isPBC).reshape(3)


def pairs_differences_to_point(point, coords, basis, isPBC, ncores=1):
    """pairs_distances.pyx:580-617 -- point-coords[i] under the boundary conditions -> (N,3)."""
    p = L.as_array(point, "point", _F32, 1).reshape(1, 3)
    return _run(p, None, None, _coords(coords), _basis(basis), isPBC, +1, True)[:, :, 0].copy()


def pairs_differences_to_indexcoords(atomIndex, coords, basis, isPBC, allAtoms=True, ncores=1):
    """pairs_distances.pyx:626-666 -- coords[atomIndex]-coords[i] under the boundary conditions."""
    idx = np.array([atomIndex], dtype=_I32)
    start = None if allAtoms else np.array([atomIndex], dtype=np.int64)
    return _run(None, idx, start, _coords(coords), _basis(basis), isPBC, +1, True)[:, :, 0].copy()


def pairs_differences_to_multi_points(points, coords, basis, isPBC, ncores=1):
    """pairs_distances.pyx:675-718 -- points is (3,k) (column t is a point) -> (N,3,k)."""
    pts = L.as_array(points, "points", _F32, 2)
    if pts.shape[0] != 3:
        raise ValueError("points must be (3,k)")
    return _run(np.ascontiguousarray(pts.T), None, None, _coords(coords), _basis(basis), isPBC, +1, True)


def pairs_differences_to_multi_indexcoords(indexes, coords, basis, isPBC, allAtoms=True, ncores=1):
    """pairs_distances.pyx:727-777 -> (N,3,k)."""
    idx = L.as_array(indexes, "indexes", _I32, 1)
    start = None if allAtoms else idx.astype(np.int64)
    return _run(None, idx, start, _coords(coords), _basis(basis), isPBC, +1, True)


def point_to_point_distance(point1, point2, basis, isPBC, ncores=1):
    """pairs_distances.pyx:786-818 -> scalar distance."""
    p1 = L.as_array(point1, "point1", _F32, 1).reshape(1, 3)
    p2 = L.as_array(point2, "point2", _F32, 1).reshape(1, 3)
    return float(_run(p1, None, None, p2, _basis(basis), isPBC, +1, False)[0, 0])


def pairs_distances_to_point(point, coords, basis, isPBC, ncores=1):
    """pairs_distances.pyx:827-866 -> (N,)."""
    p = L.as_array(point, "point", _F32, 1).reshape(1, 3)
    return _run(p, None, None, _coords(coords), _basis(basis), isPBC, +1, False)[:, 0].copy()


def pairs_distances_to_indexcoords(atomIndex, coords, basis, isPBC, allAtoms=True, ncores=1):
    """pairs_distances.pyx:874-916 -> (N,); rows below atomIndex are 0 when allAtoms=False."""
    idx = np.array([atomIndex], dtype=_I32)
    start = None if allAtoms else np.array([atomIndex], dtype=np.int64)
    return _run(None, idx, start, _coords(coords), _basis(basis), isPBC, +1, False)[:, 0].copy()


def pairs_distances_to_multi_points(points, coords, basis, isPBC, ncores=1):
    """pairs_distances.pyx:924-966 -- points is (3,k) -> (N,k)."""
    pts = L.as_array(points, "points", _F32, 2)
    if pts.shape[0] != 3:
        raise ValueError("points must be (3,k)")
    return _run(np.ascontiguousarray(pts.T), None, None, _coords(coords), _basis(basis), isPBC, +1, False)


def pairs_distances_to_multi_indexcoords(indexes, coords, basis, isPBC, allAtoms=True, ncores=1):
    """pairs_distances.pyx:975-1024 -> (N,k)."""
    idx = L.as_array(indexes, "indexes", _I32, 1)
    start = None if allAtoms else idx.astype(np.int64)
    return _run(None, idx, start, _coords(coords), _basis(basis), isPBC, +1, False)
