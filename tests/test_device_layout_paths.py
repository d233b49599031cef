"""The device-side ordering (csrc/devlayout.cu) has two implementations of each stage: one CTA per k-d node / the
multi-CTA split for large nodes, and rank counting / the segmented bitonic sort for full leaves.  Both pairs define the
same total order, so the store they build is the same record for record -- checked here on the layout itself
(frmc_debug_device_layout), each combination in its own process (the switches are read once)."""
import hashlib
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import ctypes, hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from fullrmc_b200 import _lib as L, synthetic
lib = L.load_library()
out = []
systems = [synthetic.cfg5(), synthetic.cfg4()]
rng = np.random.default_rng(5)
n = 70001                                             # non-periodic, negative coordinates, a ragged last leaf per element
box = (rng.normal(0, 30, (n, 3))).astype(np.float32)
box[:50] = 0.0; box[50:100] = -0.0                    # ties and both zeros
el = rng.integers(0, 3, n).astype(np.int32)
mol = (np.arange(n) // 3).astype(np.int32)
cases = [(s.boxCoords, s.moleculeIndex, s.elementIndex, s.numberOfElements, 1) for s in systems] + [(box, mol, el, 3, 0)]
# one element, sizes around the thresholds of the multi-CTA split (20 000 points per node, chunks of 4 096) and of the leaves
for m in (1000, 19999, 20000, 20001, 20481, 24576, 65536, 81921):
    b1 = rng.random((m, 3)).astype(np.float32)
    cases.append((b1, np.arange(m, dtype=np.int32), np.zeros(m, np.int32), 1, 1))
for coords, mol, el, nEl, pbc in cases:
    n = coords.shape[0]
    cap = n + 256 * nEl
    orig = np.zeros(cap, np.uint32)
    npad = ctypes.c_int64(0)
    seg = np.zeros(nEl + 1, np.int64)
    rc = lib.frmc_debug_device_layout(0, n, L.ptr(np.ascontiguousarray(coords), L.c_f32p), L.ptr(np.ascontiguousarray(mol), L.c_i32p),
                                      L.ptr(np.ascontiguousarray(el), L.c_i32p), nEl, pbc, cap,
                                      orig.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(npad), L.ptr(seg, L.c_i64p))
    assert rc == 0, L.last_error()
    real = orig[:npad.value]
    real = real[real != 0xFFFFFFFF]
    assert np.array_equal(np.sort(real), np.arange(n, dtype=np.uint32))          # a permutation of the atoms
    out.append(hashlib.sha256(orig[:npad.value].tobytes()).hexdigest())
print(" ".join(out))
''' % ROOT


def _layout_digests(**env):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().split()


def test_wide_split_and_bitonic_leaves_build_the_same_store():
    base = _layout_digests(FRMC_WIDE_SPLIT="0", FRMC_LEAF_FAST="0")
    assert len(base) == 11
    assert _layout_digests(FRMC_WIDE_SPLIT="1", FRMC_LEAF_FAST="0") == base
    assert _layout_digests(FRMC_WIDE_SPLIT="0", FRMC_LEAF_FAST="1") == base
    assert _layout_digests() == base
