"""CPU tests: the oracle restatements against the reference's golden vectors and, when
oracle/_ref is built, against the reference's own compiled kernels on fresh inputs."""
import os

import numpy as np
import pytest

import cases as C
from oracle import epilogue as ep

CASES = C.make_cases()
HKEYS = ("basis", "isPBC", "moleculeIndex", "elementIndex", "numberOfElements", "minDistance", "maxDistance", "bin",
         "histSize")


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "stateless.npz"))


def _kw(case):
    return {k: case[k] for k in HKEYS}


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_inputs_reproducible(case, golden):
    """the seeded generators still produce the inputs the golden outputs were made from"""
    for k, v in case.items():
        if k != "name":
            assert np.array_equal(np.asarray(v), golden["%s/in/%s" % (case["name"], k)]), k


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_c_oracle_full_histogram_matches_golden(case, golden, orc):
    hi, he, ov = orc.full_pairs_histograms_coords(boxCoords=case["boxCoords"], return_overflow=True, **_kw(case))
    assert np.array_equal(hi, golden[case["name"] + "/full/intra"])
    assert np.array_equal(he, golden[case["name"] + "/full/inter"])
    assert ov == 0
    # threaded variant of the oracle (used as the multi-core CPU baseline) is identical
    hi2, he2 = orc.full_pairs_histograms_coords(boxCoords=case["boxCoords"], ncores=4, **_kw(case))
    assert np.array_equal(hi, hi2) and np.array_equal(he, he2)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_c_oracle_multiple_and_distances_match_golden(case, golden, orc):
    nm = case["name"]
    idx = golden[nm + "/multi/indexes"]
    for allAtoms in (True, False):
        hi, he = orc.multiple_pairs_histograms_coords(indexes=idx, boxCoords=case["boxCoords"], allAtoms=allAtoms, **_kw(case))
        assert np.array_equal(hi, golden["%s/multi/%d/intra" % (nm, allAtoms)])
        assert np.array_equal(he, golden["%s/multi/%d/inter" % (nm, allAtoms)])
    a = int(golden[nm + "/dist/atom"])
    d = orc.pairs_distances_to_indexcoords(a, case["boxCoords"], case["basis"], case["isPBC"])
    assert np.array_equal(d, golden[nm + "/dist/all"])
    df = orc.pairs_differences_to_indexcoords(a, case["boxCoords"], case["basis"], case["isPBC"])
    assert np.array_equal(df, golden[nm + "/diff/all"])
    p = golden[nm + "/point"]
    assert np.array_equal(orc.pairs_distances_to_point(p, case["boxCoords"], case["basis"], case["isPBC"]),
                          golden[nm + "/dist/point"])
    assert np.array_equal(orc.pairs_differences_to_point(p, case["boxCoords"], case["basis"], case["isPBC"]),
                          golden[nm + "/diff/point"])


def test_c_oracle_dists_variants_consistent(orc):
    """*_dists functions fed with the oracle's own distances equal the *_coords functions"""
    case = CASES[1]
    n = case["boxCoords"].shape[0]
    idx = np.array([3, 4, 5, 100], dtype=np.int32)
    dist = np.stack([orc.pairs_distances_to_indexcoords(int(a), case["boxCoords"], case["basis"], case["isPBC"])
                     for a in idx], axis=1)
    kw = _kw(case)
    kw.pop("basis"); kw.pop("isPBC")
    a = orc.multiple_pairs_histograms_dists(idx, dist, **kw)
    b = orc.multiple_pairs_histograms_coords(indexes=idx, boxCoords=case["boxCoords"], **_kw(case))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert dist.shape == (n, 4)


def test_reciprocal_matches_golden(golden, orc):
    r, G, q = golden["recip/r"], golden["recip/G"], golden["recip/q"]
    assert np.array_equal(orc.Gr_to_sq(r, G, q), golden["recip/Gr_to_sq"])
    g = (G * np.float32(0.1) + np.float32(1)).astype(np.float32)
    assert np.array_equal(orc.gr_to_sq(r, g, q, 0.085), golden["recip/gr_to_sq"])


def test_c_oracle_matches_live_reference(ref_modules, orc):
    """fresh random inputs straight against the compiled reference (skipped where oracle/_ref is absent)"""
    if ref_modules is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    pd, ph, rs = ref_modules
    rng = np.random.default_rng(99)
    for trial in range(4):
        n = int(rng.integers(50, 700))
        nEl = int(rng.integers(1, 5))
        box = (rng.random((n, 3), dtype=np.float32) * np.float32(2.5) - np.float32(0.7)).astype(np.float32)
        basis = (np.eye(3) * 20 + rng.normal(0, 2.0, (3, 3))).astype(np.float32)
        kw = dict(basis=basis, isPBC=bool(trial % 2 == 0), moleculeIndex=(np.arange(n) // 3).astype(np.int32),
                  elementIndex=rng.integers(0, nEl, n).astype(np.int32), numberOfElements=nEl,
                  minDistance=np.float32(0.3), maxDistance=np.float32(9.7), bin=np.float32(0.04), histSize=235)
        a = ph.full_pairs_histograms_coords(boxCoords=box, **kw)
        b = orc.full_pairs_histograms_coords(boxCoords=box, **kw)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_round_half_away_from_zero(orc):
    """pairs_distances.pyx:31-32: a fractional difference of exactly +0.5 maps to image -0.5"""
    box = np.array([[0.75, 0.0, 0.0], [0.25, 0.0, 0.0]], dtype=np.float32)
    basis = np.diag([10.0, 10.0, 10.0]).astype(np.float32)
    d = orc.pairs_differences_to_indexcoords(0, box, basis, True)
    assert d[1, 0] == np.float32(-5.0)
    d = orc.pairs_differences_to_indexcoords(1, box, basis, True)
    assert d[0, 0] == np.float32(5.0)


def test_numpy_summation_orders_restated_exactly():
    """the two numpy reduction orders the CUDA epilogue mirrors"""
    rng = np.random.default_rng(5)
    for n in (1, 7, 8, 9, 127, 128, 129, 417, 1000, 1999, 2000, 4099):
        a = (rng.standard_normal(n) ** 2).astype(np.float32)
        assert np.add.reduce(a) == ep.numpy_pairwise_sum(a), n
    G = rng.standard_normal(300).astype(np.float32)
    M = rng.standard_normal((300, 77)).astype(np.float32)
    assert np.array_equal(np.sum(G.reshape((-1, 1)) * M, axis=0), ep.sequential_Sq(G, M))


def test_m_minus_f_identity(orc):
    """sym(M-F) counts every pair that touches the group exactly once"""
    case = CASES[1]
    rng = np.random.default_rng(3)
    idx = C.group_for(case, rng)
    kw = _kw(case)
    di, de = ep.move_delta((orc.multiple_pairs_histograms_coords, orc.full_pairs_histograms_coords), idx,
                           case["boxCoords"], kw["basis"], kw["isPBC"], kw["moleculeIndex"], kw["elementIndex"],
                           kw["numberOfElements"], kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
    sym = lambda h: h + h.transpose(1, 0, 2)
    # brute force: full histogram minus full histogram of the system without the group
    keep = np.setdiff1d(np.arange(case["boxCoords"].shape[0]), idx)
    fi, fe = orc.full_pairs_histograms_coords(boxCoords=case["boxCoords"], **kw)
    kw2 = dict(kw, moleculeIndex=kw["moleculeIndex"][keep], elementIndex=kw["elementIndex"][keep])
    ri, re_ = orc.full_pairs_histograms_coords(boxCoords=case["boxCoords"][keep], **kw2)
    assert np.array_equal(sym(di), sym(fi - ri))
    assert np.array_equal(sym(de), sym(fe - re_))
