"""The documented limits of the device path (include/fullrmc_b200.h: FRMC_MAX_ELEMENTS 16, FRMC_MAX_GROUP 64,
FRMC_MAX_GRIDS 4, FRMC_MAX_MODELS 8): AT the limit the results equal the oracle's, ONE PAST it the
call fails loudly (never a silent truncation)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


def _system(n, nEl, seed, molecule_size=1):
    from fullrmc_b200 import synthetic
    basis = np.array([[30, 0, 0], [3, 29, 0], [-2, 4, 31]], dtype=F32)
    return synthetic.random_system(n, seed, basis, n_elements=nEl, molecule_size=molecule_size)


class _Raw(object):
    """a bare system for element counts beyond the synthetic generator's element names"""

    def __init__(self, n, nEl, seed, molecule_size=1):
        rng = np.random.default_rng(seed)
        self.basis = np.array([[30, 0, 0], [3, 29, 0], [-2, 4, 31]], dtype=F32)
        self.boxCoords = rng.random((n, 3)).astype(F32)
        self.elementIndex = rng.integers(0, nEl, n).astype(np.int32)
        self.elementIndex[:nEl] = np.arange(nEl)                      # every element present
        self.moleculeIndex = (np.arange(n) // molecule_size).astype(np.int32)


def _hist_kw(s, nEl, hs=150, bin=0.1):
    return dict(basis=s.basis, isPBC=True, moleculeIndex=s.moleculeIndex, elementIndex=s.elementIndex, numberOfElements=nEl,
                minDistance=F32(0.0), maxDistance=F32(hs * bin), bin=F32(bin), histSize=hs)


def test_sixteen_elements_full_histogram_and_rows(orc):
    from fullrmc_b200.Core import pairs_histograms as ph
    s = _Raw(3000, 16, 2, molecule_size=3)
    kw = _hist_kw(s, 16)
    hi, he = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    wi, we = orc.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    assert np.array_equal(hi, wi) and np.array_equal(he, we) and he.sum() > 0
    idx = np.arange(10, 74, dtype=np.int32)
    hi, he = ph.multiple_pairs_histograms_coords(indexes=idx, boxCoords=s.boxCoords, allAtoms=True, **kw)
    wi, we = orc.multiple_pairs_histograms_coords(indexes=idx, boxCoords=s.boxCoords, allAtoms=True, **kw)
    assert np.array_equal(hi, wi) and np.array_equal(he, we)


def test_seventeen_elements_are_refused():
    from fullrmc_b200.Core import pairs_histograms as ph
    from fullrmc_b200.store import DeviceStore
    s = _Raw(400, 17, 3)
    with pytest.raises(Exception) as e:
        ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **_hist_kw(s, 17))
    assert "numberOfElements" in str(e.value)
    with pytest.raises(Exception) as e:
        DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 17)
    assert "numberOfElements" in str(e.value) or "element" in str(e.value).lower()


def test_group_of_sixty_four_atoms_and_one_more(orc):
    """a 64-atom group through propose / accept: the running histograms equal a recount; 65 atoms are refused"""
    from fullrmc_b200 import synthetic
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    s = _system(2500, 4, 5, molecule_size=2)
    kw = _hist_kw(s, 4)
    rng = np.random.default_rng(1)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 4) as store:
        gi = store.add_grid(kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
        grid = synthetic.RGrid(0.0, 0.1, 150)
        store.add_model(gi, ModelSpec("PDF", experimental=np.zeros(grid.hs, F32), elements=s.elements, n_per_element=s.numberOfAtomsPerElement,
                                      weighting=s.weighting, volume=s.volume, rho0=s.numberDensity, shell_centers=grid.shellCenters,
                                      shell_volumes=grid.shellVolumes))
        store.compute_data()
        box = s.boxCoords.copy()
        for first in (100, 130):                                        # the second group overlaps the first
            idx = np.arange(first, first + 64, dtype=np.int32)
            moved = (box[idx] + rng.normal(0, 0.01, (64, 3))).astype(F32)
            store.propose(idx, moved)
            store.accept()
            box[idx] = moved
        intra, inter = store.export_data(0)
        wi, we = orc.full_pairs_histograms_coords(boxCoords=box, **kw)
        sym = lambda a: a + a.transpose(1, 0, 2)
        assert np.array_equal(sym(np.asarray(intra, np.int64)), sym(wi.astype(np.int64)))
        assert np.array_equal(sym(np.asarray(inter, np.int64)), sym(we.astype(np.int64)))
        idx = np.arange(0, 65, dtype=np.int32)
        with pytest.raises(Exception) as e:
            store.propose(idx, box[idx])
        assert "group size" in str(e.value)
        # the failed call left nothing staged: the store still works
        store.propose(idx[:3], box[idx[:3]]); store.reject()


def test_four_grids_eight_models_and_one_more():
    from fullrmc_b200.model import ModelSpec
    from fullrmc_b200.store import DeviceStore
    from fullrmc_b200 import synthetic
    s = _system(1500, 2, 7)
    with DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 2) as store:
        grids = [synthetic.RGrid(0.0, 0.05 + 0.01 * i, 100 + 10 * i) for i in range(4)]
        gids = [store.add_grid(g.minDistance, g.maxDistance, g.bin, g.hs) for g in grids]
        with pytest.raises(Exception) as e:
            store.add_grid(0.0, 5.0, 0.05, 100)
        assert "grids" in str(e.value)
        for m in range(8):
            g = grids[m % 4]
            common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume,
                          rho0=s.numberDensity, shell_centers=g.shellCenters, shell_volumes=g.shellVolumes)
            store.add_model(gids[m % 4], ModelSpec("PDF" if m % 2 == 0 else "PCF", experimental=np.zeros(g.hs, F32), **common))
        with pytest.raises(Exception) as e:
            store.add_model(gids[0], ModelSpec("PDF", experimental=np.zeros(grids[0].hs, F32), **common))
        assert "models" in str(e.value)
        chi2 = np.array(store.compute_data())
        assert chi2.shape[0] == 8 and np.all(np.isfinite(chi2))
        # two models on the same grid and of the same kind agree with each other (0 and 4, 1 and 5, ...)
        assert chi2[0] == chi2[4] and chi2[1] == chi2[5] and chi2[2] == chi2[6] and chi2[3] == chi2[7]
        idx = np.array([5], np.int32)
        after = np.array(store.propose(idx, (s.boxCoords[idx] + F32(0.01)).astype(F32)))
        assert after.shape[0] == 8 and after[0] == after[4] and after[3] == after[7]
        store.reject()


def test_histogram_too_large_for_shared_memory_is_refused():
    from fullrmc_b200.Core import pairs_histograms as ph
    s = _system(600, 2, 9)
    kw = _hist_kw(s, 2, hs=60000, bin=0.0002)
    with pytest.raises(Exception) as e:
        ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    assert "histSize" in str(e.value) or "shared memory" in str(e.value)
