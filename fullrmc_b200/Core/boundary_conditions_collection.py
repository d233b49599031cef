"""Drop-in for the function of ``fullrmc.Core.boundary_conditions_collection`` on the move path
(Extensions/boundary_conditions_collection.pyx:88-110): ``transform_coordinates``, which Engine.py:3223 calls on every
generated move to turn moved real coordinates into box coordinates.  Same name, argument names and error behaviour
(positional-only like the reference's ``always_allow_keywords(False)`` is NOT enforced: keywords are accepted too).
"""
import numpy as np

from .. import _lib as L


def transform_coordinates(transMatrix, coords):
    """(N,3) float32 = coords . transMatrix, float32 products and sums in the reference's order."""
    m = L.as_array(transMatrix, "transMatrix", np.float32, 2)
    c = L.as_array(coords, "coords", np.float32, 2)
    out = np.empty((c.shape[0], 3), dtype=np.float32)
    if c.shape[0]:
        L.check(L.load_library().frmc_transform_coordinates(L.device_index(), L.ptr(m, L.c_f32p), L.ptr(c, L.c_f32p), c.shape[0],
                                                            L.ptr(out, L.c_f32p)), "transform_coordinates")
    return out
