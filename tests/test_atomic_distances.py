"""Distance-constraint kernels (Extensions/atomic_distances.pyx).

CPU: the oracle restatement (oracle/pairhist_oracle.c: orc_multiple_atomic_distances_coords) against the golden
outputs of the reference's own compiled functions (tests/gen_golden_atomic_distances.py).
GPU: fullrmc_b200.Core.atomic_distances against the same golden vectors and against the oracle on larger systems.
Bar: counts AND float32 distance sums bit-identical (the sums depend on the reference's loop order)."""
import os

import numpy as np
import pytest

from gen_golden_atomic_distances import FLAG_SETS

KEYS = ("nintra", "dintra", "ninter", "dinter")


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "atomic_distances.npz"))


def _system(g, name):
    return dict(boxCoords=g[name + "/boxCoords"], basis=g[name + "/basis"], isPBC=bool(g[name + "/isPBC"]),
                moleculeIndex=g[name + "/moleculeIndex"], elementIndex=g[name + "/elementIndex"],
                numberOfElements=int(g[name + "/numberOfElements"]), lowerLimit=g[name + "/lowerLimit"],
                upperLimit=g[name + "/upperLimit"])


def _compare(module, g):
    assert int(g["n_flag_sets"]) == len(FLAG_SETS)
    for name in [str(x) for x in g["names"]]:
        kw = _system(g, name)
        idx = g[name + "/indexes"]
        for fi, flags in enumerate(FLAG_SETS):
            for allAtoms in (True, False):
                r = module.multiple_atomic_distances_coords(indexes=idx, allAtoms=allAtoms, **kw, **flags)
                for key, a in zip(KEYS, r):
                    ref = g["%s/multiple/%d/%d/%s" % (name, fi, int(allAtoms), key)]
                    assert a.dtype == ref.dtype and a.shape == ref.shape
                    assert np.array_equal(a, ref), (name, flags, allAtoms, key)
            r = module.full_atomic_distances_coords(**kw, **flags)
            for key, a in zip(KEYS, r):
                assert np.array_equal(a, g["%s/full/%d/%s" % (name, fi, key)]), (name, flags, key)


def test_oracle_matches_reference_golden(golden, orc):
    _compare(orc, golden)


@pytest.mark.gpu
def test_device_matches_reference_golden(golden):
    from fullrmc_b200.Core import atomic_distances
    _compare(atomic_distances, golden)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nT,pbc", [(20000, 3, True), (15000, 2, False)])
def test_device_matches_oracle_on_larger_systems(n, nT, pbc, orc):
    """minimum-approach windows on a dense system: many rows, hits in every type pair, group moves"""
    from fullrmc_b200.Core import atomic_distances as ad
    rng = np.random.default_rng(n)
    basis = np.array([[52, 0, 0], [6, 50, 0], [-4, 7, 49]], np.float32) if pbc else np.eye(3, dtype=np.float32)
    box = rng.random((n, 3)).astype(np.float32) if pbc else (rng.random((n, 3)) * 50.0).astype(np.float32)
    el = rng.integers(0, nT, n).astype(np.int32)
    mol = (np.arange(n) // 3).astype(np.int32)
    lo = np.zeros((nT, nT, 1), np.float32)
    up = (1.2 + rng.random((nT, nT, 1))).astype(np.float32); up = ((up + up.transpose(1, 0, 2)) / 2).astype(np.float32)
    kw = dict(boxCoords=box, basis=basis, isPBC=pbc, moleculeIndex=mol, elementIndex=el, numberOfElements=nT, lowerLimit=lo, upperLimit=up)
    from fullrmc_b200 import _lib
    for flags in (dict(intraMolecular=False), dict(reduceDistanceToUpper=True, intraMolecular=False), dict()):
        got = ad.full_atomic_distances_coords(**kw, **flags)            # k-d ordered store, block pairs out of reach culled
        ref = orc.full_atomic_distances_coords(**kw, **flags)
        for key, a, b in zip(KEYS, got, ref):
            assert np.array_equal(a, b), (flags, key)
        old = _lib.set_block_culling(False)                            # the plain rows sweep over the same input
        try:
            plain = ad.full_atomic_distances_coords(**kw, **flags)
        finally:
            _lib.set_block_culling(old)
        for key, a, b in zip(KEYS, plain, ref):
            assert np.array_equal(a, b), (flags, key, "plain")
        assert int(ref[2].sum()) > 100                      # the case really has close contacts
        idx = rng.integers(0, n, 40).astype(np.int32)
        got = ad.multiple_atomic_distances_coords(indexes=idx, **kw, **flags)
        ref = orc.multiple_atomic_distances_coords(idx, **kw, **flags)
        for key, a, b in zip(KEYS, got, ref):
            assert np.array_equal(a, b), (flags, key)


@pytest.mark.gpu
def test_argument_checks_like_the_reference():
    from fullrmc_b200.Core import atomic_distances as ad
    n, nT = 10, 2
    box = np.zeros((n, 3), np.float32); basis = np.eye(3, dtype=np.float32)
    mol = np.zeros(n, np.int32); el = np.zeros(n, np.int32)
    lim = np.zeros((nT, nT, 1), np.float32)
    with pytest.raises(TypeError):
        ad.full_atomic_distances_coords(None, basis, True, mol, el, nT, lim, lim)
    with pytest.raises(ValueError):
        ad.full_atomic_distances_coords(box.astype(np.float64), basis, True, mol, el, nT, lim, lim)
    with pytest.raises(AssertionError):
        ad.full_atomic_distances_coords(box, basis, True, mol, el, nT, np.zeros((3, 3, 1), np.float32), lim)
    r = ad.full_atomic_distances_coords(box, basis, True, mol, el, nT, lim, lim + 1)     # all atoms coincide: d = 0 in [0, 1)
    assert int(r[0].sum()) == n * (n - 1) // 2 and float(r[1].sum()) == 0.0


def test_pair_elements_stats_equals_the_compiled_reference(ref_modules):
    """host-side prefix-count form of atomic_distances.pyx:632-672 against the reference's O(N^2) loop (no GPU involved)"""
    import importlib
    if ref_modules is None:
        pytest.skip("oracle/_ref not built")
    ref = importlib.import_module("fullrmc.Core.atomic_distances")
    from fullrmc_b200.Core import atomic_distances as ad
    rng = np.random.default_rng(0)
    for n, nT, msize in ((1, 2, 1), (2, 2, 1), (500, 3, 13), (777, 4, 1), (600, 2, 600), (900, 5, 7)):
        el = rng.integers(0, nT, n).astype(np.int32)
        mol = (rng.permutation(n) // msize).astype(np.int32)              # molecules scattered over the index range
        want = ref.pair_elements_stats(elementIndex=el, moleculeIndex=mol, numberOfElements=nT)
        got = ad.pair_elements_stats(elementIndex=el, moleculeIndex=mol, numberOfElements=nT)
        assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
        assert got[0].dtype == np.int32 and got[0].shape == (nT, nT, 1)
