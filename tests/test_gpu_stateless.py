"""GPU parity tests of the stateless drop-in modules (fullrmc_b200.Core.*) against the
golden vectors of the reference and against the oracle on larger seeded inputs.
Bar: bit-exact float32 histograms, distances and difference vectors."""
import os

import numpy as np
import pytest

import cases as C

pytestmark = pytest.mark.gpu

CASES = C.make_cases()
HKEYS = ("basis", "isPBC", "moleculeIndex", "elementIndex", "numberOfElements", "minDistance", "maxDistance", "bin",
         "histSize")


def _kw(case):
    return {k: case[k] for k in HKEYS}


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "stateless.npz"))


@pytest.fixture(scope="module")
def ph():
    from fullrmc_b200.Core import pairs_histograms
    return pairs_histograms


@pytest.fixture(scope="module")
def pdm():
    from fullrmc_b200.Core import pairs_distances
    return pairs_distances


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_full_histogram_matches_reference_golden(case, golden, ph):
    hi, he = ph.full_pairs_histograms_coords(boxCoords=case["boxCoords"], **_kw(case))
    assert hi.dtype == np.float32 and hi.shape == (case["numberOfElements"],) * 2 + (case["histSize"],)
    assert np.array_equal(hi, golden[case["name"] + "/full/intra"])
    assert np.array_equal(he, golden[case["name"] + "/full/inter"])
    assert ph.LAST_EDGE_OVERFLOW == 0


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_multiple_histograms_match_reference_golden(case, golden, ph):
    nm = case["name"]
    idx = golden[nm + "/multi/indexes"]
    for allAtoms in (True, False):
        hi, he = ph.multiple_pairs_histograms_coords(indexes=idx, boxCoords=case["boxCoords"], allAtoms=allAtoms, **_kw(case))
        assert np.array_equal(hi, golden["%s/multi/%d/intra" % (nm, allAtoms)])
        assert np.array_equal(he, golden["%s/multi/%d/inter" % (nm, allAtoms)])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_distances_and_differences_match_reference_golden(case, golden, pdm):
    nm = case["name"]
    a = int(golden[nm + "/dist/atom"])
    box, basis, pbc = case["boxCoords"], case["basis"], case["isPBC"]
    assert np.array_equal(pdm.pairs_distances_to_indexcoords(atomIndex=a, coords=box, basis=basis, isPBC=pbc),
                          golden[nm + "/dist/all"])
    assert np.array_equal(pdm.pairs_differences_to_indexcoords(atomIndex=a, coords=box, basis=basis, isPBC=pbc),
                          golden[nm + "/diff/all"])
    p = golden[nm + "/point"]
    assert np.array_equal(pdm.pairs_distances_to_point(point=p, coords=box, basis=basis, isPBC=pbc), golden[nm + "/dist/point"])
    assert np.array_equal(pdm.pairs_differences_to_point(point=p, coords=box, basis=basis, isPBC=pbc), golden[nm + "/diff/point"])
    # allAtoms=False: rows from the atom index on are identical, rows before are zero here
    d = pdm.pairs_distances_to_indexcoords(atomIndex=a, coords=box, basis=basis, isPBC=pbc, allAtoms=False)
    assert np.array_equal(d[a:], golden[nm + "/dist/all"][a:]) and not d[:a].any()


def test_remaining_distance_functions_against_oracle(pdm, orc):
    case = CASES[1]
    box, basis = case["boxCoords"], case["basis"]
    n = box.shape[0]
    idx = np.array([0, 17, n - 1], dtype=np.int32)
    for pbc in (True, False):
        d = pdm.pairs_distances_to_multi_indexcoords(indexes=idx, coords=box, basis=basis, isPBC=pbc)
        df = pdm.pairs_differences_to_multi_indexcoords(indexes=idx, coords=box, basis=basis, isPBC=pbc)
        assert d.shape == (n, 3) and df.shape == (n, 3, 3)
        for t, a in enumerate(idx):
            assert np.array_equal(d[:, t], orc.pairs_distances_to_indexcoords(int(a), box, basis, pbc))
            assert np.array_equal(df[:, :, t], orc.pairs_differences_to_indexcoords(int(a), box, basis, pbc))
        pts = np.ascontiguousarray((box[idx] + np.float32(0.21)).T)          # (3,k)
        d = pdm.pairs_distances_to_multi_points(points=pts, coords=box, basis=basis, isPBC=pbc)
        df = pdm.pairs_differences_to_multi_points(points=pts, coords=box, basis=basis, isPBC=pbc)
        for t in range(3):
            assert np.array_equal(d[:, t], orc.pairs_distances_to_point(pts[:, t].copy(), box, basis, pbc))
            assert np.array_equal(df[:, :, t], orc.pairs_differences_to_point(pts[:, t].copy(), box, basis, pbc))
        p1, p2 = box[3].copy(), box[11].copy()
        dd = pdm.point_to_point_distance(point1=p1, point2=p2, basis=basis, isPBC=pbc)
        assert np.float32(dd) == orc.pairs_distances_to_point(p1, p2.reshape(1, 3), basis, pbc)[0]
        one = pdm.pair_difference_to_point(point1=p1, point2=p2, basis=basis, isPBC=pbc)
        assert np.array_equal(one, orc.pairs_differences_to_point(p2, p1.reshape(1, 3), basis, pbc)[0])
        ft = pdm.from_to_points_differences(pointsFrom=box[:50].copy(), pointsTo=box[50:100].copy(), basis=basis, isPBC=pbc)
        for i in range(50):
            assert np.array_equal(ft[i], orc.pairs_differences_to_point(box[50 + i].copy(), box[i:i + 1].copy(), basis, pbc)[0])


def test_dists_and_single_variants(ph, orc):
    case = CASES[2]
    kw = _kw(case)
    box = case["boxCoords"]
    idx = np.array([5, 6, 7, 300], dtype=np.int32)
    dist = np.stack([orc.pairs_distances_to_indexcoords(int(a), box, kw["basis"], kw["isPBC"]) for a in idx], axis=1)
    kd = {k: kw[k] for k in kw if k not in ("basis", "isPBC")}
    for allAtoms in (True, False):
        a = ph.multiple_pairs_histograms_dists(indexes=idx, distances=dist, allAtoms=allAtoms, **kd)
        b = orc.multiple_pairs_histograms_dists(idx, dist, allAtoms=allAtoms, **kd)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    n = box.shape[0]
    full = np.stack([orc.pairs_distances_to_indexcoords(a, box, kw["basis"], kw["isPBC"]) for a in range(60)], axis=1)
    sub = dict(kd, moleculeIndex=kd["moleculeIndex"][:60].copy(), elementIndex=kd["elementIndex"][:60].copy())
    a = ph.full_pairs_histograms_dists(distances=np.ascontiguousarray(full[:60]), **sub)
    b = orc.full_pairs_histograms_dists(np.ascontiguousarray(full[:60]), **sub)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # single_pairs_histograms updates in place, twice in a row accumulates
    nEl, hs = kw["numberOfElements"], kw["histSize"]
    hi = np.zeros((nEl, nEl, hs), np.float32); he = np.zeros((nEl, nEl, hs), np.float32)
    ri = np.zeros((nEl, nEl, hs), np.float32); re_ = np.zeros((nEl, nEl, hs), np.float32)
    for a_idx in (5, 300):
        col = np.ascontiguousarray(dist[:, list(idx).index(a_idx)])
        ph.single_pairs_histograms(a_idx, col, kd["moleculeIndex"], kd["elementIndex"], hi, he,
                                   kw["minDistance"], kw["maxDistance"], kw["bin"])
        orc.single_pairs_histograms(a_idx, col, kd["moleculeIndex"], kd["elementIndex"], ri, re_,
                                    kw["minDistance"], kw["maxDistance"], kw["bin"])
    assert hi.sum() > 0 and np.array_equal(hi, ri) and np.array_equal(he, re_)
    assert n == dist.shape[0]


@pytest.mark.parametrize("name,n,nEl,basis,spread,molsize", [
    ("ortho_R4", 40000, 5, np.diag([74.0, 73.0, 75.0]), None, 1),
    ("tri_R4_molecular", 36000, 3, np.array([[70, 0, 0], [9, 68, 0], [-7, 13, 66]]), None, 13),
    ("tri_R4_unwrapped", 34000, 2, np.array([[70, 0, 0], [9, 68, 0], [-7, 13, 66]]), 0.8, 1),
])
def test_full_histogram_multi_tile_against_oracle(name, n, nEl, basis, spread, molsize, ph, orc):
    """systems large enough for the R=4 register-tiled kernel, several I-tiles and J-chunks per element"""
    from fullrmc_b200 import synthetic
    s = synthetic.random_system(n, 11, np.asarray(basis, dtype=np.float32), n_elements=nEl, molecule_size=molsize, spread=spread)
    kw = dict(s.hist_kwargs(), minDistance=np.float32(0.0), maxDistance=np.float32(20.0), bin=np.float32(0.02), histSize=1000)
    hi, he = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    ri, re_, ov = orc.full_pairs_histograms_coords(boxCoords=s.boxCoords, ncores=orc.max_threads(), return_overflow=True, **kw)
    assert np.array_equal(hi, ri) and np.array_equal(he, re_)
    assert ph.LAST_EDGE_OVERFLOW == ov
    # every in-range unordered pair counted exactly once
    assert hi.sum(dtype=np.float64) + he.sum(dtype=np.float64) == ri.sum(dtype=np.float64) + re_.sum(dtype=np.float64)


_TRI = np.array([[120, 0, 0], [17, 115, 0], [-13, 21, 118]], dtype=np.float32)


@pytest.mark.parametrize("name,n,nEl,basis,isPBC,spread,rmax", [
    ("ortho_fast", 50000, 2, np.diag([120.0, 118.0, 122.0]), True, None, 9.0),
    ("tri_fast", 44000, 3, _TRI, True, None, 9.0),
    ("ortho_general_unwrapped", 42000, 2, np.diag([120.0, 118.0, 122.0]), True, 2.3, 9.0),
    ("tri_general_unwrapped", 40000, 2, _TRI, True, 1.7, 9.0),
    ("infinite_boundaries", 40000, 2, np.diag([120.0, 118.0, 122.0]), False, None, 9.0),
    ("rmax_beyond_half_box", 20000, 2, np.diag([40.0, 41.0, 39.0]), True, None, 30.0),
])
def test_block_culling_never_changes_the_histogram(name, n, nEl, basis, isPBC, spread, rmax, ph, orc):
    """Morton-ordered blocks whose bounding boxes are farther apart than maxDistance are skipped; the
    histogram must equal the plain O(N^2) sweep and the oracle in every geometry mode (periodic seam,
    unwrapped coordinates, skewed cell, no PBC, maxDistance larger than half the box)."""
    from fullrmc_b200 import _lib, synthetic
    s = synthetic.random_system(n, 23, np.asarray(basis, dtype=np.float32), n_elements=nEl, molecule_size=7, isPBC=isPBC,
                                spread=spread)
    hs = 300
    kw = dict(s.hist_kwargs(), minDistance=np.float32(0.5), maxDistance=np.float32(rmax),
              bin=np.float32((rmax - 0.5) / hs), histSize=hs)
    culled = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    old = _lib.set_block_culling(False)
    try:
        plain = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    finally:
        _lib.set_block_culling(old)
    ri, re_ = orc.full_pairs_histograms_coords(boxCoords=s.boxCoords, ncores=orc.max_threads(), **kw)
    assert ri.sum() + re_.sum() > 0
    assert np.array_equal(plain[0], ri) and np.array_equal(plain[1], re_)
    assert np.array_equal(culled[0], ri) and np.array_equal(culled[1], re_)


def _adversarial_systems():
    """coordinate sets chosen to stress the bounding-box bound: the periodic seam, exact 0/1, unwrapped
    offsets, degenerate (flat / linear) boxes, a strongly skewed cell, large non-periodic coordinates"""
    rng = np.random.default_rng(77)
    n = 5200
    out = []
    ortho = np.diag([60.0, 55.0, 65.0]).astype(np.float32)
    skew = np.array([[60, 0, 0], [48, 30, 0], [-35, 25, 40]], dtype=np.float32)
    # two tight clusters on either side of the seam in x, one in the middle
    c = np.concatenate([rng.normal([0.01, 0.3, 0.3], 0.012, (n // 3, 3)), rng.normal([0.99, 0.3, 0.3], 0.012, (n // 3, 3)),
                        rng.normal([0.5, 0.7, 0.7], 0.05, (n - 2 * (n // 3), 3))]).astype(np.float32)
    out.append(("seam_clusters", c, ortho, True))
    # exact grid values including 0, 1, -0.0 and whole-cell offsets
    g = (rng.integers(0, 33, (n, 3)) / 32.0).astype(np.float32)
    g[::7] += np.float32(3.0); g[1::7] -= np.float32(2.0); g[2::11] = np.float32(-0.0)
    out.append(("lattice_points_and_offsets", g, ortho, True))
    # all atoms in one plane / on one line: zero-width boxes
    flat = rng.random((n, 3)).astype(np.float32); flat[:, 2] = np.float32(0.25)
    out.append(("flat_sheet", flat, ortho, True))
    line = rng.random((n, 3)).astype(np.float32); line[:, 1] = np.float32(0.5); line[:, 2] = np.float32(0.999999)
    out.append(("line", line, skew, True))
    out.append(("skewed_cell", rng.random((n, 3)).astype(np.float32), skew, True))
    out.append(("skewed_cell_unwrapped", (rng.random((n, 3)) * 5.0 - 2.5).astype(np.float32), skew, True))
    far = (rng.random((n, 3)) * 90.0 + np.array([480.0, -520.0, 3.0])).astype(np.float32)
    out.append(("non_periodic_large_coordinates", far, np.eye(3, dtype=np.float32), False))
    return out


@pytest.mark.parametrize("name,coords,basis,isPBC", _adversarial_systems(), ids=[a[0] for a in _adversarial_systems()])
@pytest.mark.parametrize("rmax", [2.5, 7.0, 26.0])
def test_block_culling_adversarial_geometries(name, coords, basis, isPBC, rmax, ph, orc):
    n = coords.shape[0]
    rng = np.random.default_rng(5)
    el = rng.integers(0, 2, n).astype(np.int32)
    mol = (np.arange(n) // 3).astype(np.int32)
    hs = 50
    kw = dict(basis=basis, isPBC=isPBC, moleculeIndex=mol, elementIndex=el, numberOfElements=2,
              minDistance=np.float32(0.0), maxDistance=np.float32(rmax), bin=np.float32(rmax / hs), histSize=hs)
    hi, he = ph.full_pairs_histograms_coords(boxCoords=coords, **kw)
    ri, re_, ov = orc.full_pairs_histograms_coords(boxCoords=coords, ncores=orc.max_threads(), return_overflow=True, **kw)
    assert np.array_equal(hi, ri) and np.array_equal(he, re_)
    assert ph.LAST_EDGE_OVERFLOW == ov


def test_block_culling_skips_most_of_a_sparse_system():
    """the store reports how many distance evaluations the last compute_data made"""
    from fullrmc_b200 import _lib, synthetic
    from fullrmc_b200.store import DeviceStore
    s = synthetic.cfg5(120000, seed=3)
    st = DeviceStore(s.boxCoords, s.basis, s.isPBC, s.moleculeIndex, s.elementIndex, s.numberOfElements)
    g = st.add_grid(0.0, 8.0, 0.02, 400)
    st.compute_data()
    swept = st.swept_pairs
    intra, inter = st.export_data(g)
    old = _lib.set_block_culling(False)
    try:
        st.compute_data()
        swept_all = st.swept_pairs
        intra2, inter2 = st.export_data(g)
    finally:
        _lib.set_block_culling(old)
        st.close()
    n = s.boxCoords.shape[0]
    assert swept_all >= n * (n - 1) // 2
    assert 0 < swept < 0.25 * swept_all
    assert np.array_equal(intra, intra2) and np.array_equal(inter, inter2)


def test_chunk_level_culling_only_removes_evaluations(orc):
    """the finer level of the culling (8-record chunks inside the 32-record units): fewer distance evaluations, the same
    histogram -- against the unit-level sweep and against the oracle, in a triclinic and a non-periodic system"""
    from fullrmc_b200 import _lib, synthetic
    from fullrmc_b200.store import DeviceStore
    tri = np.array([[60, 0, 0], [9, 58, 0], [-7, 11, 57]], dtype=np.float32)
    for s, pbc in ((synthetic.random_system(30000, 5, tri, n_elements=3, molecule_size=3), True),
                   (synthetic.random_system(20000, 6, np.eye(3, dtype=np.float32), n_elements=2, isPBC=False, spread=40.0), False)):
        st = DeviceStore(s.boxCoords, s.basis, pbc, s.moleculeIndex, s.elementIndex, s.numberOfElements)
        g = st.add_grid(0.5, 9.0, 0.05, 170)
        st.compute_data()
        swept = st.swept_pairs
        intra, inter = st.export_data(g)
        old = _lib.set_chunk_culling(False)
        try:
            st.compute_data()
            swept_units = st.swept_pairs
            intra2, inter2 = st.export_data(g)
        finally:
            _lib.set_chunk_culling(old)
            st.close()
        assert 0 < swept < swept_units
        assert np.array_equal(intra, intra2) and np.array_equal(inter, inter2)
        wi, we = orc.full_pairs_histograms_coords(boxCoords=s.boxCoords, basis=s.basis, isPBC=pbc, moleculeIndex=s.moleculeIndex,
                                                  elementIndex=s.elementIndex, numberOfElements=s.numberOfElements,
                                                  minDistance=np.float32(0.5), maxDistance=np.float32(9.0), bin=np.float32(0.05), histSize=170)
        assert np.array_equal(intra, wi) and np.array_equal(inter, we)


def test_full_histogram_shards_sum_to_whole(ph):
    """the multi-GPU decomposition: per-shard partial histograms add up to the single-call result"""
    from fullrmc_b200 import synthetic
    s = synthetic.random_system(9000, 12, np.diag([45.0, 45.0, 45.0]).astype(np.float32), n_elements=3)
    kw = dict(s.hist_kwargs(), minDistance=np.float32(0.0), maxDistance=np.float32(15.0), bin=np.float32(0.05), histSize=300)
    whole = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    for nshards in (2, 8):
        acc_i = np.zeros_like(whole[0]); acc_e = np.zeros_like(whole[1])
        for shard in range(nshards):
            hi, he = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, _shard=shard, _nshards=nshards, **kw)
            acc_i += hi; acc_e += he
        assert np.array_equal(acc_i, whole[0]) and np.array_equal(acc_e, whole[1])


def test_edge_bin_overflow_is_counted_like_the_oracle(ph, orc):
    """a grid whose last fp32 quotient rounds up to histSize: the reference writes out of bounds
    (boundscheck False); oracle and GPU drop the event and report the same count"""
    rng = np.random.default_rng(8)
    hs = 37
    b = np.float32(0.1)
    rmax = np.float32(np.float32(hs) * b) + np.float32(3e-6)    # max slightly above hs*bin
    n = 3000
    box = rng.random((n, 3), dtype=np.float32)
    basis = np.diag([9.0, 9.0, 9.0]).astype(np.float32)
    kw = dict(basis=basis, isPBC=True, moleculeIndex=np.arange(n, dtype=np.int32), elementIndex=np.zeros(n, np.int32),
              numberOfElements=1, minDistance=np.float32(0.0), maxDistance=rmax, bin=b, histSize=hs)
    ri, re_, ov = orc.full_pairs_histograms_coords(boxCoords=box, return_overflow=True, **kw)
    hi, he = ph.full_pairs_histograms_coords(boxCoords=box, **kw)
    assert np.array_equal(hi, ri) and np.array_equal(he, re_)
    assert ph.LAST_EDGE_OVERFLOW == ov
    idx = np.arange(0, n, 7, dtype=np.int32)
    ri, re_, ov = orc.multiple_pairs_histograms_coords(indexes=idx, boxCoords=box, return_overflow=True, **kw)
    hi, he = ph.multiple_pairs_histograms_coords(indexes=idx, boxCoords=box, **kw)
    assert np.array_equal(hi, ri) and np.array_equal(he, re_) and ph.LAST_EDGE_OVERFLOW == ov


def test_edge_bin_spill_mode_reproduces_the_reference_write(ph, orc, ref_modules):
    """with set_edge_spill(True) the in-array spill of the reference's unchecked write is reproduced:
    compared with the oracle in spill mode and, when built, with the compiled reference itself"""
    import fullrmc_b200
    rng = np.random.default_rng(8)
    hs = 37
    b = np.float32(0.1)
    rmax = np.float32(np.float32(hs) * b) + np.float32(3e-6)
    n = 3000
    box = rng.random((n, 3), dtype=np.float32)
    basis = np.diag([9.0, 9.0, 9.0]).astype(np.float32)
    el = np.zeros(n, np.int32); el[n // 2:] = 1        # element 1 only at the high indexes: slabs [0,0],[0,1],[1,1] used, [1,0] empty
    kw = dict(basis=basis, isPBC=True, moleculeIndex=np.arange(n, dtype=np.int32), elementIndex=el,
              numberOfElements=2, minDistance=np.float32(0.0), maxDistance=rmax, bin=b, histSize=hs)
    previous = fullrmc_b200.set_edge_spill(True)
    orc.set_emulate_spill(True)
    try:
        ri, re_, ov = orc.full_pairs_histograms_coords(boxCoords=box, return_overflow=True, **kw)
        hi, he = ph.full_pairs_histograms_coords(boxCoords=box, **kw)
        assert ov > 0 and ph.LAST_EDGE_OVERFLOW == ov
        assert np.array_equal(hi, ri) and np.array_equal(he, re_)
        idx = np.arange(0, n, 5, dtype=np.int32)
        ri2, re2, ov2 = orc.multiple_pairs_histograms_coords(indexes=idx, boxCoords=box, return_overflow=True, **kw)
        hi2, he2 = ph.multiple_pairs_histograms_coords(indexes=idx, boxCoords=box, **kw)
        assert np.array_equal(hi2, ri2) and np.array_equal(he2, re2) and ph.LAST_EDGE_OVERFLOW == ov2
    finally:
        fullrmc_b200.set_edge_spill(previous)
        orc.set_emulate_spill(False)
    # the spill lands where the reference itself writes (all spills stay inside the array here:
    # slab [1,1] pairs would leave it, so only compare when none of its events overflowed)
    if ref_modules is not None:
        fi, fe = ref_modules[1].full_pairs_histograms_coords(boxCoords=box, **kw)
        same = np.array_equal(fe[:1], he[:1]) and np.array_equal(fe[1, 0], he[1, 0])
        assert same, "in-array spill differs from the compiled reference"


def test_empty_and_degenerate_inputs(ph):
    basis = np.eye(3, dtype=np.float32)
    kw = dict(basis=basis, isPBC=True, numberOfElements=2, minDistance=np.float32(0.0), maxDistance=np.float32(1.0),
              bin=np.float32(0.1), histSize=10)
    box = np.random.default_rng(0).random((5, 3), dtype=np.float32)
    mol = np.zeros(5, np.int32); el = np.zeros(5, np.int32)
    hi, he = ph.multiple_pairs_histograms_coords(indexes=np.zeros(0, np.int32), boxCoords=box, moleculeIndex=mol, elementIndex=el, **kw)
    assert not hi.any() and not he.any()
    with pytest.raises(ValueError):
        ph.multiple_pairs_histograms_coords(indexes=np.array([7], np.int32), boxCoords=box, moleculeIndex=mol, elementIndex=el, **kw)
    with pytest.raises(ValueError):
        ph.full_pairs_histograms_coords(boxCoords=box, moleculeIndex=mol, elementIndex=np.full(5, 3, np.int32), **kw)


def test_reciprocal_space_functions(golden, orc):
    from fullrmc_b200.Core import reciprocal_space as rs
    r, G, q = golden["recip/r"], golden["recip/G"], golden["recip/q"]
    sq = rs.Gr_to_sq(distances=r, Gr=G, qrange=q)
    # double-precision sine on the device is not glibc's: tolerance 1e-6 relative (north_star), normwise
    assert np.max(np.abs(sq - golden["recip/Gr_to_sq"])) <= 1e-6 * np.max(np.abs(golden["recip/Gr_to_sq"]))
    g = (G * np.float32(0.1) + np.float32(1)).astype(np.float32)
    sq = rs.gr_to_sq(distances=r, gr=g, qrange=q, rho=np.float32(0.085))
    assert np.max(np.abs(sq - golden["recip/gr_to_sq"])) <= 1e-6 * np.max(np.abs(golden["recip/gr_to_sq"]))
    # sq_to_Gr: no oracle in the reference (the function raises); check against the documented formula in float64
    Gr = rs.sq_to_Gr(qValues=q, rValues=r[:50].copy(), sq=golden["recip/Gr_to_sq"])
    dq = float(q[1] - q[0])
    want = np.array([(2 / np.pi) * np.sum(q.astype(np.float64) * (golden["recip/Gr_to_sq"].astype(np.float64) - 1) *
                                          np.sin(q.astype(np.float64) * float(rr)) * dq) for rr in r[:50]])
    assert np.max(np.abs(Gr - want)) <= 5e-5 * max(1.0, np.max(np.abs(want)))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_device_and_host_layout_give_the_same_histogram(case, ph, orc):
    """the stateless full histogram orders the atoms on the device by default (csrc/devlayout.cu); the host k-d ordering
    (frmc_set_device_layout(0)) and the oracle must give the identical arrays"""
    from fullrmc_b200 import _lib
    want = orc.full_pairs_histograms_coords(boxCoords=case["boxCoords"], ncores=orc.max_threads(), **_kw(case))
    for on in (True, False):
        previous = _lib.set_device_layout(on)
        try:
            got = ph.full_pairs_histograms_coords(boxCoords=case["boxCoords"], **_kw(case))
        finally:
            _lib.set_device_layout(previous)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), "device layout %s" % on
