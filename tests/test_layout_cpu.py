"""CPU tests of the host side of the device store layout (csrc/fullhist.cu:build_layout through the host-only
entry point frmc_debug_layout): the element-sorted, k-d ordered record order the full-histogram kernel culls on.
No device is needed; the kernels that consume the layout are tested under -m gpu."""
import ctypes

import numpy as np
import pytest

from fullrmc_b200 import _lib as L

PAD = 0xFFFFFFFF


def _layout(coords, el, nEl, isPBC=True, mol=None, device=False):
    lib = L.load_library()
    n = coords.shape[0]
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    el = np.ascontiguousarray(el, dtype=np.int32)
    mol = np.arange(n, dtype=np.int32) if mol is None else np.ascontiguousarray(mol, dtype=np.int32)
    cap = n + 256 * nEl
    orig = np.empty(cap, dtype=np.uint32)
    npad = ctypes.c_int64(0)
    seg = np.zeros(nEl + 1, dtype=np.int64)
    args = (n, L.ptr(coords, L.c_f32p), L.ptr(mol, L.c_i32p), L.ptr(el, L.c_i32p), nEl, int(isPBC), cap,
            orig.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(npad), L.ptr(seg, L.c_i64p))
    if device:                                                   # the layout the device builds (csrc/devlayout.cu)
        L.check(lib.frmc_debug_device_layout(L.device_index(), *args), "debug_device_layout")
    else:
        L.check(lib.frmc_debug_layout(*args), "debug_layout")
    return orig[:npad.value], seg


SHAPES = [(1, 1, True, 0.0), (300, 3, True, 0.0), (5000, 2, False, 0.0), (70000, 5, True, 0.0), (40000, 4, True, 2.5)]


@pytest.mark.gpu
@pytest.mark.parametrize("n,nEl,pbc,spread", SHAPES + [(1025, 1, True, 0.0), (200000, 1, True, 0.0), (33, 16, False, 0.0)])
def test_device_layout_is_an_element_sorted_permutation(n, nEl, pbc, spread):
    _check_permutation(n, nEl, pbc, spread, device=True)


@pytest.mark.parametrize("n,nEl,pbc,spread", SHAPES)
def test_layout_is_an_element_sorted_permutation(n, nEl, pbc, spread):
    _check_permutation(n, nEl, pbc, spread, device=False)


def _check_permutation(n, nEl, pbc, spread, device):
    rng = np.random.default_rng(n + nEl)
    coords = (rng.random((n, 3)) * (1 + 2 * spread) - spread).astype(np.float32)
    if not pbc:
        coords = (coords * 80.0 - 13.0).astype(np.float32)
    el = rng.integers(0, nEl, n).astype(np.int32)
    orig, seg = _layout(coords, el, nEl, pbc, device=device)
    assert orig.shape[0] == seg[-1] and orig.shape[0] % 256 == 0
    real = orig[orig != PAD]
    assert np.array_equal(np.sort(real), np.arange(n, dtype=np.uint32))          # every atom exactly once
    for e in range(nEl):
        block = orig[seg[e]:seg[e + 1]]
        cnt = int(np.sum(el == e))
        assert (seg[e + 1] - seg[e]) == (cnt + 255) // 256 * 256
        assert np.all(block[:cnt] != PAD) and np.all(block[cnt:] == PAD)          # padding only at the end of a segment
        assert np.all(el[block[:cnt]] == e)
    # deterministic
    orig2, _ = _layout(coords, el, nEl, pbc, device=device)
    assert np.array_equal(orig, orig2)


@pytest.mark.gpu
def test_device_kd_order_makes_compact_blocks():
    _check_compact(device=True)


def test_kd_order_makes_compact_blocks():
    """uniform points: aligned runs of 256 (and 32) records must be boxes of about the ideal volume; a random order
    of the same points would give boxes spanning the whole cell"""
    _check_compact(device=False)


def _check_compact(device):
    rng = np.random.default_rng(5)
    n = 120000
    coords = rng.random((n, 3)).astype(np.float32)
    el = rng.integers(0, 2, n).astype(np.int32)
    orig, seg = _layout(coords, el, 2, True, device=device)
    for blk, slack in ((256, 2.0), (32, 3.0)):
        vols = []
        for e in range(2):
            cnt = int(np.sum(el == e))
            idx = orig[seg[e]:seg[e] + cnt // blk * blk].reshape(-1, blk)
            pts = coords[idx]                                   # [blocks, blk, 3]
            vols.append(np.prod(pts.max(1) - pts.min(1), axis=1))
        vols = np.concatenate(vols)
        ideal = blk / (n / 2.0)                                  # volume that holds blk points of one element on average
        assert np.median(vols) < slack * ideal, (blk, np.median(vols), ideal)
        assert vols.max() < 6.0 * ideal


def test_unwrapped_coordinates_are_ordered_on_their_periodic_image():
    """atoms that differ by whole cells belong to the same place: the order must not depend on the integer offsets"""
    rng = np.random.default_rng(9)
    n = 30000
    base = rng.random((n, 3)).astype(np.float32)
    shifted = (base + rng.integers(-2, 3, (n, 3)).astype(np.float32)).astype(np.float32)
    frac = shifted - np.floor(shifted)
    keep = np.all(np.abs(frac - base) < 1e-6, axis=1)           # drop the few atoms whose fp32 fractional part moved
    el = rng.integers(0, 3, n).astype(np.int32)
    o1, _ = _layout(frac[keep], el[keep], 3, True)
    o2, _ = _layout(shifted[keep], el[keep], 3, True)
    assert np.array_equal(o1, o2)
