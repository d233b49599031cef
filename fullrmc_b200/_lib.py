"""ctypes binding of libfullrmc_b200.so (the C ABI declared in include/fullrmc_b200.h).

The library is loaded lazily; a missing library is a hard error (no fallback path).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libfullrmc_b200.so"
_lib = None


class FullrmcB200Error(RuntimeError):
    """Raised when the CUDA library reports an error through the C ABI."""


c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u64p = ctypes.POINTER(ctypes.c_uint64)


class ModelDesc(ctypes.Structure):
    """Mirror of ``struct frmc_model_desc`` (include/fullrmc_b200.h)."""
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("n_pairs", ctypes.c_int32),
        ("pair_a", c_i32p),
        ("pair_b", c_i32p),
        ("pair_w", c_f32p),
        ("pair_D", c_f32p),
        ("shell_volumes", c_f32p),
        ("prefactor", c_f32p),
        ("shape", c_f32p),
        ("scale", ctypes.c_float),
        ("n_out", ctypes.c_int32),
        ("experimental", c_f32p),
        ("data_weights", c_f32p),
        ("gr2sq", c_f32p),
        ("sq_exact", ctypes.c_int32),
    ]


class AmputationDesc(ctypes.Structure):
    """Mirror of ``struct frmc_amputation_desc`` (include/fullrmc_b200.h)."""
    _fields_ = [("pair_w", c_f32p), ("pair_D", c_f32p), ("prefactor", c_f32p)]


# every symbol include/fullrmc_b200.h declares: name -> (restype, argtypes)
_I, _I64, _F, _VP = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
SIGNATURES = {
    "frmc_last_error": (ctypes.c_char_p, []),
    "frmc_version": (ctypes.c_char_p, []),
    "frmc_device_count": (_I, []),
    "frmc_set_edge_spill": (_I, [_I]),
    "frmc_set_block_culling": (_I, [_I]),
    "frmc_set_chunk_culling": (_I, [_I]),
    "frmc_ctx_set_timing": (_I, [_I, _I]),
    "frmc_ctx_kernel_ms": (_I, [_I, ctypes.POINTER(ctypes.c_double)]),
    "frmc_set_device_layout": (_I, [_I]),
    "frmc_launch_count": (ctypes.c_uint64, []),
    "frmc_points_to_coords": (_I, [_I, c_f32p, c_i32p, c_i64p, _I64, c_f32p, _I64, c_f32p, _I, _I, _I, c_f32p]),
    "frmc_from_to_points_differences": (_I, [_I, c_f32p, c_f32p, _I64, c_f32p, _I, c_f32p]),
    "frmc_multiple_pairs_histograms_coords": (_I, [_I, c_i32p, _I64, c_f32p, _I64, c_f32p, _I, c_i32p, c_i32p, _I,
                                                   _F, _F, _F, _I, _I, c_f32p, c_f32p, c_u64p]),
    "frmc_full_pairs_histograms_coords": (_I, [_I, c_f32p, _I64, c_f32p, _I, c_i32p, c_i32p, _I, _F, _F, _F, _I,
                                               _I, _I, c_f32p, c_f32p, c_u64p]),
    "frmc_multiple_atomic_distances_coords": (_I, [_I, c_i32p, _I64, c_f32p, _I64, c_f32p, _I, c_i32p, c_i32p, _I, c_f32p, c_f32p, _I, _I,
                                                   c_i32p, c_f32p, c_i32p, c_f32p]),
    "frmc_full_atomic_distances_coords": (_I, [_I, c_f32p, _I64, c_f32p, _I, c_i32p, c_i32p, _I, c_f32p, c_f32p, _I,
                                               c_i32p, c_f32p, c_i32p, c_f32p]),
    "frmc_coordination_counts": (_I, [_I, c_f32p, _I64, c_f32p, _I, c_f32p, _I64, _I64, c_i32p, c_i32p, c_i32p, c_f32p, c_f32p,
                                      _I64, c_i64p, c_i32p, _I64, c_i32p]),
    "frmc_debug_work_items": (_I, [_I64, c_i32p, _I, _I, _I, _I, c_i64p, c_i64p]),
    "frmc_full_pairs_histograms_coords_multi": (_I, [_I, ctypes.POINTER(ctypes.c_int), c_f32p, _I64, c_f32p, _I, c_i32p, c_i32p, _I, _F, _F, _F, _I,
                                                     c_f32p, c_f32p, ctypes.POINTER(ctypes.c_uint64)]),
    "frmc_multi_reduce_path": (ctypes.c_char_p, []),
    "frmc_shape_function": (_I, [_I, c_f32p, c_f32p, _I, _I, _I, c_i32p, c_i32p, ctypes.POINTER(ctypes.c_double), c_f32p, c_f32p,
                                 ctypes.c_double, c_f32p, _I, c_f32p, _I, c_f32p]),
    "frmc_debug_layout": (_I, [_I64, c_f32p, c_i32p, c_i32p, _I, _I, _I64, ctypes.POINTER(ctypes.c_uint32), c_i64p, c_i64p]),
    "frmc_debug_device_layout": (_I, [_I, _I64, c_f32p, c_i32p, c_i32p, _I, _I, _I64, ctypes.POINTER(ctypes.c_uint32), c_i64p, c_i64p]),
    "frmc_multiple_pairs_histograms_dists": (_I, [_I, c_i32p, _I64, c_f32p, _I64, c_i32p, c_i32p, _I, _F, _F, _F, _I,
                                                  _I, c_f32p, c_f32p, c_u64p]),
    "frmc_single_pairs_histograms": (_I, [_I, ctypes.c_int32, c_f32p, _I64, _I64, c_i32p, c_i32p, _I, _I, c_f32p,
                                          c_f32p, _F, _F, _F, _I, c_u64p]),
    "frmc_Gr_to_sq": (_I, [_I, c_f32p, c_f32p, _I64, c_f32p, _I64, c_f32p]),
    "frmc_gr_to_sq": (_I, [_I, c_f32p, c_f32p, _I64, c_f32p, _I64, _F, c_f32p]),
    "frmc_sq_to_Gr": (_I, [_I, c_f32p, c_f32p, c_f32p, _I64, _I64, c_f32p]),
    "frmc_store_create": (_VP, [_I, _I64, c_f32p, c_f32p, _I, c_i32p, c_i32p, _I]),
    "frmc_store_destroy": (None, [_VP]),
    "frmc_store_set_coords": (_I, [_VP, c_f32p, c_f32p]),
    "frmc_store_get_coords": (_I, [_VP, c_f32p]),
    "frmc_store_stream": (_VP, [_VP]),
    "frmc_grid_add": (_I, [_VP, _F, _F, _F, _I]),
    "frmc_model_add": (_I, [_VP, _I, ctypes.POINTER(ModelDesc)]),
    "frmc_model_set_scale": (_I, [_VP, _I, _F]),
    "frmc_model_set_shape": (_I, [_VP, _I, c_f32p]),
    "frmc_model_set_window": (_I, [_VP, _I, c_f32p, _I]),
    "frmc_model_set_multiframe_prior": (_I, [_VP, _I, c_f32p, _F]),
    "frmc_model_set_adjust": (_I, [_VP, _I, _I, _F, _F]),
    "frmc_model_get_scale": (_I, [_VP, _I, c_f32p, c_f32p]),
    "frmc_store_set_accepted": (_I, [_VP, ctypes.c_uint64]),
    "frmc_store_set_persistent": (_I, [_VP, _I]),
    "frmc_store_persistent_stats": (_I, [_VP, c_u64p, c_u64p]),
    "frmc_compute_data": (_I, [_VP, c_f32p]),
    "frmc_compute_data_shard": (_I, [_VP, _I, _I]),
    "frmc_grid_counts_ptr": (_VP, [_VP, _I, c_i64p]),
    "frmc_finalize_data": (_I, [_VP, c_f32p]),
    "frmc_propose": (_I, [_VP, c_i32p, _I, c_f32p, c_f32p]),
    "frmc_accept": (_I, [_VP]),
    "frmc_reject": (_I, [_VP]),
    # array arguments as plain addresses: the tight loop passes arr.__array_interface__["data"][0] (1 us) instead of
    # building a typed ctypes pointer per call (4 us each)
    "frmc_step": (_I, [_VP, _I, _VP, _I, _VP, c_f32p]),
    "frmc_propose_amputation": (_I, [_VP, ctypes.c_int32, ctypes.POINTER(AmputationDesc), _I, c_f32p]),
    "frmc_accept_amputation": (_I, [_VP]),
    "frmc_reject_amputation": (_I, [_VP]),
    "frmc_model_set_constants": (_I, [_VP, _I, c_f32p, c_f32p, c_f32p]),
    "frmc_store_n_atoms": (_I64, [_VP]),
    "frmc_import_data": (_I, [_VP, _I, c_f32p, c_f32p]),
    "frmc_transform_coordinates": (_I, [_I, c_f32p, c_f32p, _I64, c_f32p]),
    "frmc_store_set_real_coords": (_I, [_VP, c_f32p, c_f32p]),
    "frmc_store_get_real_coords": (_I, [_VP, c_f32p]),
    "frmc_store_set_groups": (_I, [_VP, _I, c_i32p, c_i32p]),
    "frmc_run_generated": (_I, [_VP, _I, ctypes.c_uint64, ctypes.c_uint64, _F, _F, c_f32p, _F, c_f32p, c_f32p, c_i32p, c_i32p, c_f32p,
                                ctypes.POINTER(ctypes.c_double)]),
    "frmc_store_distance_add": (_I, [_VP, c_i32p, _I, c_f32p, c_f32p, _I]),
    # array arguments as plain addresses (see frmc_step)
    "frmc_store_distance_move": (_I, [_VP, _I, _VP, _I, _VP, _VP, _VP]),
    "frmc_store_move_atoms": (_I, [_VP, c_i32p, _I, c_f32p]),
    "frmc_store_coordination_add": (_I, [_VP, _I, c_i64p, c_i32p, c_i64p, c_i32p, c_f32p, c_f32p]),
    "frmc_store_coordination_move": (_I, [_VP, _I, _VP, _I, _VP, _VP]),
    "frmc_run_batch": (_I, [_VP, _I, c_i32p, c_i32p, c_f32p, c_f32p, _F, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p,
                            ctypes.POINTER(ctypes.c_double)]),
    "frmc_store_batch_stats": (_I, [_VP, c_u64p, c_u64p, c_u64p]),
    "frmc_store_committed_chi2": (_I, [_VP, c_f32p]),
    "frmc_store_batch_stamps": (_I, [_VP, c_i64p, _I]),
    "frmc_store_replay_proposal": (_I, [_VP, _I, ctypes.POINTER(ctypes.c_double)]),
    "frmc_export_data": (_I, [_VP, _I, c_f32p, c_f32p]),
    "frmc_export_total": (_I, [_VP, _I, _I, c_f32p]),
    "frmc_store_edge_overflow": (ctypes.c_uint64, [_VP]),
    "frmc_store_swept_pairs": (ctypes.c_uint64, [_VP]),
    "frmc_store_debug_stamps": (_I, [_VP, c_i64p, _I]),
    "frmc_store_set_timing": (_I, [_VP, _I]),
    "frmc_store_get_timing": (_I, [_VP, _I, ctypes.POINTER(ctypes.c_double), c_u64p]),
}


def library_path():
    return os.environ.get("FULLRMC_B200_LIB", os.path.join(_HERE, "lib", _LIB_NAME))


def load_library():
    """Load libfullrmc_b200.so and bind every declared symbol.  Raises RuntimeError when
    the library has not been built (there is deliberately no other backend)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("fullrmc_b200: CUDA library %s not found -- build it with "
                           "fullrmc_b200/csrc/build.sh (python __graft_entry__.py build). "
                           "There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    msg = load_library().frmc_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


_ERR_CLASSES = {-1: ValueError, -4: RuntimeError, -5: ValueError}


def check(rc, what):
    """Turn a negative return code into the exception the reference would have raised:
    ValueError for bad arguments (Cython buffer validation raises ValueError), RuntimeError
    (FullrmcB200Error) for CUDA failures."""
    if rc >= 0:
        return rc
    msg = "%s: %s" % (what, last_error())
    cls = _ERR_CLASSES.get(rc, FullrmcB200Error)
    raise cls(msg)


def set_edge_spill(on):
    """Edge-bin policy for grids and stateless calls created from now on (see include/fullrmc_b200.h):
    False (default) drops pairs whose fp32 bin index rounds up to histSize, True reproduces the
    reference's in-array spill of that unchecked write.  Returns the previous setting."""
    return bool(load_library().frmc_set_edge_spill(int(bool(on))))


def set_block_culling(on):
    """Full-histogram block culling (default on; results are identical either way).  Returns the previous setting."""
    return bool(load_library().frmc_set_block_culling(int(bool(on))))


def set_chunk_culling(on):
    """Chunk-level culling inside the 32-record units of the full histogram (default on; identical results either way).
    Returns the previous setting."""
    return bool(load_library().frmc_set_chunk_culling(int(bool(on))))


def kernel_ms_of(call, device=None):
    """Run ``call()`` (a stateless drop-in function) with the library's kernel timer on; returns (result, device ms of
    the call's dominant kernel)."""
    lib = load_library()
    dev = device_index() if device is None else int(device)
    check(lib.frmc_ctx_set_timing(dev, 1), "ctx_set_timing")
    try:
        out = call()
        ms = ctypes.c_double(0.0)
        check(lib.frmc_ctx_kernel_ms(dev, ctypes.byref(ms)), "ctx_kernel_ms")
    finally:
        lib.frmc_ctx_set_timing(dev, 0)
    return out, float(ms.value)


def set_device_layout(on):
    """Stateless full histogram: order the caller's atoms on the device (default) or on the host cores.  Identical
    histograms either way.  Returns the previous setting."""
    return bool(load_library().frmc_set_device_layout(int(bool(on))))


def device_index():
    """Device used by the stateless drop-in modules: $FULLRMC_B200_DEVICE, else LOCAL_RANK, else 0."""
    for key in ("FULLRMC_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(key)
        if v is not None and v != "":
            return int(v)
    return 0


def device_list():
    """Devices the stateless full histogram spreads over: $FULLRMC_B200_DEVICES = "0,1,2,3" or "all" (one Python
    process, several GPUs: frmc_full_pairs_histograms_coords_multi); unset = the single device of device_index()."""
    v = os.environ.get("FULLRMC_B200_DEVICES", "").strip()
    if not v:
        return [device_index()]
    if v.lower() == "all":
        return list(range(int(load_library().frmc_device_count())))
    return [int(x) for x in v.split(",") if x.strip() != ""]


# ---------------------------------------------------------------- argument validation
# Mirrors what Cython's typed-buffer arguments do in the reference wrappers
# (e.g. pairs_histograms.pyx:150-162): None -> TypeError, wrong dtype / ndim -> ValueError.
def as_array(a, name, dtype, ndim, allow_none=False):
    if a is None:
        if allow_none:
            return None
        raise TypeError("Argument '%s' must not be None" % name)
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
    if a.dtype != dtype:
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'" % (np.dtype(dtype).name, a.dtype.name))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    return np.ascontiguousarray(a)


def ptr(a, ctype):
    return a.ctypes.data_as(ctype) if a is not None else ctypes.cast(None, ctype)
