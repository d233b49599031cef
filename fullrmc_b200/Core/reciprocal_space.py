"""Drop-in for ``fullrmc.Core.reciprocal_space`` (reference: Extensions/reciprocal_space.pyx).

``gr_to_sq`` / ``Gr_to_sq`` reproduce the reference's mixed arithmetic (double-precision
sine term rounded to fp32, fp32 accumulation in r order).  ``sq_to_Gr`` implements the
documented formula; the reference function itself raises TypeError (it calls libc ``sin``
on an ndarray, reciprocal_space.pyx:143), so it has no oracle.
"""
import numpy as np

from .. import _lib as L

_F32 = np.float32


def gr_to_sq(distances, gr, qrange, rho):
    """reciprocal_space.pyx:42-73 -- S(q) = 1 + 4*pi*rho * sum_r dr*r*sin(q r)/q*(g(r)-1)."""
    lib = L.load_library()
    r = L.as_array(distances, "distances", _F32, 1)
    g = L.as_array(gr, "gr", _F32, 1)
    q = L.as_array(qrange, "qrange", _F32, 1)
    if g.shape != r.shape:
        raise ValueError("distances and gr must have the same length")
    sq = np.empty(q.shape[0], dtype=_F32)
    rc = lib.frmc_gr_to_sq(L.device_index(), L.ptr(r, L.c_f32p), L.ptr(g, L.c_f32p), r.shape[0],
                           L.ptr(q, L.c_f32p), q.shape[0], float(_F32(rho)), L.ptr(sq, L.c_f32p))
    L.check(rc, "gr_to_sq")
    return sq


def Gr_to_sq(distances, Gr, qrange):
    """reciprocal_space.pyx:82-109 -- S(q) = 1 + sum_r dr*sin(q r)/q*G(r)."""
    lib = L.load_library()
    r = L.as_array(distances, "distances", _F32, 1)
    g = L.as_array(Gr, "Gr", _F32, 1)
    q = L.as_array(qrange, "qrange", _F32, 1)
    if g.shape != r.shape:
        raise ValueError("distances and Gr must have the same length")
    sq = np.empty(q.shape[0], dtype=_F32)
    rc = lib.frmc_Gr_to_sq(L.device_index(), L.ptr(r, L.c_f32p), L.ptr(g, L.c_f32p), r.shape[0],
                           L.ptr(q, L.c_f32p), q.shape[0], L.ptr(sq, L.c_f32p))
    L.check(rc, "Gr_to_sq")
    return sq


def sq_to_Gr(qValues, rValues, sq):
    """reciprocal_space.pyx:118-145 -- G(r) = (2/pi) sum_q q (S(q)-1) sin(q r) dq."""
    lib = L.load_library()
    q = L.as_array(qValues, "qValues", _F32, 1)
    r = L.as_array(rValues, "rValues", _F32, 1)
    s = L.as_array(sq, "sq", _F32, 1)
    if s.shape != q.shape:
        raise ValueError("qValues and sq must have the same length")
    Gr = np.empty(r.shape[0], dtype=_F32)
    rc = lib.frmc_sq_to_Gr(L.device_index(), L.ptr(q, L.c_f32p), L.ptr(r, L.c_f32p), L.ptr(s, L.c_f32p),
                           q.shape[0], r.shape[0], L.ptr(Gr, L.c_f32p))
    L.check(rc, "sq_to_Gr")
    return Gr
