"""Parity AT THE SIZES bench.py measures: BASELINE.json configs[3] (cfg4: 100 000 atoms, triclinic, 5 elements)
and configs[4] (cfg5: 1 000 000 atoms, cubic).  Fixtures: tests/gen_golden_large.py.

* full_cfg4.npz / constraints_cfg4.npz -- the compiled reference and the unmodified reference constraint
  classes throughout.
* full_cfg5.npz / constraints_cfg5.npz -- full histograms from oracle/pairhist_oracle.c (pinned to the compiled
  reference; 5e11 pairs take 4.3 h through the reference's Python row loop), cross-checked against the compiled
  reference on 64 rows stored with it; the 200-move trajectory is the reference's own class code.

Bars: every histogram cell, every chi^2, every decision bit-exact.
"""
import os

import numpy as np
import pytest

import test_golden_constraints as TG

F32 = np.float32
NAMES = ["cfg4", "cfg5"]


def _have(golden_dir, *names):
    return all(os.path.exists(os.path.join(golden_dir, n)) for n in names)


def _full(golden_dir, name):
    from fullrmc_b200 import synthetic
    z = np.load(os.path.join(golden_dir, "full_%s.npz" % name))
    system = getattr(synthetic, str(z["recipe_name"]))(int(z["recipe_n"]), int(z["recipe_seed"]))
    kw = dict(system.hist_kwargs(), minDistance=F32(z["minDistance"]), maxDistance=F32(z["maxDistance"]), bin=F32(z["bin"]),
              histSize=int(z["histSize"]))
    return z, system, kw


@pytest.mark.parametrize("name", NAMES)
def test_fixture_rows_match_the_oracle(name, golden_dir, orc):
    """CPU: the 64 reference rows stored with the full histogram are what the pinned C oracle computes, and the
    full histogram holds them (every row is part of the upper triangle)"""
    if not _have(golden_dir, "full_%s.npz" % name):
        pytest.skip("fixture not generated")
    z, system, kw = _full(golden_dir, name)
    oi, oe = orc.multiple_pairs_histograms_coords(indexes=z["rows"], boxCoords=system.boxCoords, allAtoms=False, **kw)
    assert np.array_equal(oi, z["rows_intra"]) and np.array_equal(oe, z["rows_inter"])
    assert np.all(z["rows_intra"] <= z["intra"]) and np.all(z["rows_inter"] <= z["inter"])
    n = system.numberOfAtoms
    assert z["intra"].shape == (5, 5, int(z["histSize"])) and float(z["intra"].sum(dtype=np.float64) + z["inter"].sum(dtype=np.float64)) < n * (n - 1) / 2


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_full_histogram_at_benchmark_size(name, golden_dir):
    """culled, un-culled, 8-shard sum, the rows entry point and the store path against the fixture"""
    from fullrmc_b200 import _lib as fullrmc_b200
    from fullrmc_b200.Core import pairs_histograms as ph
    from fullrmc_b200.store import DeviceStore
    if not _have(golden_dir, "full_%s.npz" % name):
        pytest.skip("fixture not generated")
    z, system, kw = _full(golden_dir, name)
    hi, he = ph.full_pairs_histograms_coords(boxCoords=system.boxCoords, **kw)
    assert np.array_equal(hi, z["intra"]) and np.array_equal(he, z["inter"]), "culled sweep differs from the reference"
    assert ph.LAST_EDGE_OVERFLOW == int(z["edge_overflow"])
    previous = fullrmc_b200.set_block_culling(False)
    try:
        bi, be = ph.full_pairs_histograms_coords(boxCoords=system.boxCoords, **kw)
    finally:
        fullrmc_b200.set_block_culling(previous)
    assert np.array_equal(bi, z["intra"]) and np.array_equal(be, z["inter"]), "plain sweep differs from the reference"
    si, se = np.zeros_like(hi), np.zeros_like(he)
    for shard in range(8):
        a, b = ph.full_pairs_histograms_coords(boxCoords=system.boxCoords, _shard=shard, _nshards=8, **kw)
        si += a; se += b
    assert np.array_equal(si, z["intra"]) and np.array_equal(se, z["inter"]), "8-shard sum differs from the reference"
    ri, re_ = ph.multiple_pairs_histograms_coords(indexes=z["rows"], boxCoords=system.boxCoords, allAtoms=False, **kw)
    assert np.array_equal(ri, z["rows_intra"]) and np.array_equal(re_, z["rows_inter"]), "rows differ from the compiled reference"
    with DeviceStore(system.boxCoords, system.basis, True, system.moleculeIndex, system.elementIndex, 5) as store:
        g = store.add_grid(kw["minDistance"], kw["maxDistance"], kw["bin"], kw["histSize"])
        store.compute_data_shard(0, 1)
        di, de = store.export_data(g)
        assert np.array_equal(di, z["intra"]) and np.array_equal(de, z["inter"]), "store path differs from the reference"


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_reference_class_trajectory_at_benchmark_size(name, golden_dir):
    """frmc_propose / frmc_accept / frmc_reject through the constraint mirrors: chi^2 of all 200 moves, final data"""
    import fullrmc_b200
    from fullrmc_b200.constraints import DeviceBackend, make_device_constraint
    if not _have(golden_dir, "constraints_%s.npz" % name):
        pytest.skip("fixture not generated")
    previous = fullrmc_b200.set_edge_spill(True)
    try:
        TG._replay_on_device(name, golden_dir, DeviceBackend, make_device_constraint)
    finally:
        fullrmc_b200.set_edge_spill(previous)


def _backend(g):
    from fullrmc_b200.constraints import DeviceBackend, make_device_constraint
    elements, n_per = TG._system(g)
    backend = DeviceBackend(g["boxCoords"], g["basis"], bool(g["isPBC"]), g["moleculeIndex"], g["elementIndex"], elements,
                            n_per, g["volume"], g["numberDensity"])
    cons = []
    for ci in range(int(g["n_constraints"])):
        d = TG._constraint_desc(g, ci)
        cons.append((d, make_device_constraint(backend, d["kind"], d["experimental"], d["minDistance"], d["maxDistance"], d["bin"],
                                               int(d["histSize"]), d["shellCenters"], d["shellVolumes"], d["weighting"],
                                               dataWeights=d["dataWeights"], scaleFactor=float(d["scaleFactor"]),
                                               qValues=d.get("qValues") if d["kind"] in ("SQ", "RSQ") else None)))
    for ci, (d, c) in enumerate(cons):
        _, err = c.compute_data()
        assert F32(err) == F32(g["start_stdErr"][ci])
    return backend, cons


def _check_final(g, backend, cons):
    for ci, (d, c) in enumerate(cons):
        hi, he = backend.store.export_data(c._grid)
        assert np.array_equal(hi, d["final_intra"]) and np.array_equal(he, d["final_inter"])
        assert F32(backend.store.committed_chi2()[c._model]) == F32(d["final_stdErr"])
        assert np.array_equal(backend.store.export_total(c._model), d["final_total"])
    assert np.array_equal(backend.store.get_coords(), g["final_boxCoords"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("path", ["step", "step_persistent", "run_batch"])
def test_step_and_run_batch_replay_at_benchmark_size(name, path, golden_dir):
    """the same trajectory through frmc_step (resolve + propose in one call, launch per move and resident kernel) and
    through frmc_run_batch (the engine's rule on the device, the recorded decisions forced through the random numbers)"""
    import fullrmc_b200
    from test_gpu_batch import forced_random_numbers
    if not _have(golden_dir, "constraints_%s.npz" % name):
        pytest.skip("fixture not generated")
    g = TG._load(golden_dir, name)
    previous = fullrmc_b200.set_edge_spill(True)
    try:
        backend, cons = _backend(g)
        store = backend.store
        steps = g["steps/idx"].shape[0]
        nc = len(cons)
        want = g["steps/chi2_after"][:steps, :nc].astype(F32)
        if path == "run_batch":
            rand, total0 = forced_random_numbers(g)
            ks = g["steps/k"][:steps].astype(np.int32)
            idx = np.concatenate([g["steps/idx"][s, :ks[s]] for s in range(steps)]).astype(np.int32)
            moved = np.concatenate([g["steps/moved"][s, :ks[s]] for s in range(steps)]).astype(F32)
            out = store.run_batch(idx, moved, total0, rand, tolerance=0.5, group_sizes=ks)
            assert np.array_equal(out["decisions"] > 0, g["steps/accepted"][:steps])
            assert np.array_equal(out["chi2"], want), "chi2 differs from the reference classes"
            assert store.batch_stats()[0] >= 1
        else:
            if path == "step_persistent":
                store.set_persistent(True)
            prev = None
            for s in range(steps):
                k = int(g["steps/k"][s])
                chi = store.step(prev, g["steps/idx"][s, :k].astype(np.int32), np.ascontiguousarray(g["steps/moved"][s, :k]))
                assert np.array_equal(chi[:nc].astype(F32), want[s]), "step %d" % s
                prev = bool(g["steps/accepted"][s])
            (store.accept if prev else store.reject)()
            if path == "step_persistent":
                assert store.persistent_stats()[1] >= steps
                store.set_persistent(False)
        _check_final(g, backend, cons)
        backend.close()
    finally:
        fullrmc_b200.set_edge_spill(previous)
