"""Coordination-number counts (Extensions/atomic_coordination.pyx).

CPU: the oracle restatement (oracle/coordination.py + orc_single_atom_single_shell_* in oracle/pairhist_oracle.c)
against the golden outputs of the reference's own compiled module (tests/gen_golden_atomic_coordination.py).
GPU: fullrmc_b200.Core.atomic_coordination against the same golden vectors, against the oracle on larger systems, and
through the before/after bookkeeping of AtomicCoordinationNumberConstraint (AtomicCoordinationConstraints.py:519-577).
All comparisons are exact: the outputs are integer counts held as float32.
"""
import os

import numpy as np
import pytest

from gen_golden_atomic_coordination import definitions, run_all, unpack_lists


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "atomic_coordination.npz"))


@pytest.fixture(scope="module")
def orc_pd():
    from oracle import pairhist
    return pairhist


def _system(g, name):
    defs = (unpack_lists(g, name + "/cores"), unpack_lists(g, name + "/shells"), [np.float32(x) for x in g[name + "/lowerShells"]],
            [np.float32(x) for x in g[name + "/upperShells"]], unpack_lists(g, name + "/asCore", True), unpack_lists(g, name + "/inShell", True))
    return g[name + "/boxCoords"], g[name + "/basis"], bool(g[name + "/isPBC"]), defs, g[name + "/indexes"]


def _compare(module, pd, g):
    names = [str(x) for x in g["names"]]
    assert len(names) == 4
    for name in names:
        box, basis, pbc, defs, idx = _system(g, name)
        before = box.copy()
        res = run_all(module, pd, box, basis, pbc, defs, idx)
        assert np.array_equal(box, before)                       # inputs are borrowed, never modified
        keys = sorted(k[len(name) + 5:] for k in g.files if k.startswith(name + "/out/"))
        assert keys == sorted(res.keys())
        for key in keys:
            ref = g["%s/out/%s" % (name, key)]
            got = np.asarray(res[key])
            assert got.dtype == ref.dtype and got.shape == ref.shape, (name, key)
            assert np.array_equal(got, ref), (name, key, got, ref)
        assert bool(g[name + "/out/all_totdists_raises_TypeError"])
        assert g[name + "/out/all_coords"].sum() > 100           # the fixture is not vacuous


def test_oracle_matches_reference_golden(golden, orc_pd):
    from oracle import coordination
    _compare(coordination, orc_pd, golden)


@pytest.mark.gpu
def test_device_matches_reference_golden(golden):
    from fullrmc_b200.Core import atomic_coordination, pairs_distances
    _compare(atomic_coordination, pairs_distances, golden)


def _random_system(n, nT, pbc, ndef, seed):
    rng = np.random.default_rng(seed)
    basis = np.array([[40, 0, 0], [5, 38, 0], [-3, 6, 36]], np.float32) if pbc else np.eye(3, dtype=np.float32)
    box = (rng.random((n, 3)) * 1.6 - 0.3).astype(np.float32)     # partly outside [0,1): the general wrap is exercised
    if not pbc:
        box = (box * 36.0).astype(np.float32)
    el = rng.integers(0, nT, n).astype(np.int32)
    return rng, box, basis, definitions(rng, n, el, nT, ndef)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nT,pbc", [(6000, 3, True), (5000, 2, False)])
def test_device_matches_oracle_on_larger_systems(n, nT, pbc):
    """lists longer than one work item (2048 entries), thousands of tasks in one launch"""
    from fullrmc_b200.Core import atomic_coordination as dev
    from oracle import coordination as orc
    rng, box, basis, defs = _random_system(n, nT, pbc, 4, 77 + n)
    cores, shells, lowers, uppers, as_core, in_shell = defs
    kw = dict(boxCoords=box, basis=basis, isPBC=pbc, coresIndexes=cores, shellsIndexes=shells, lowerShells=lowers, upperShells=uppers,
              asCoreDefIdxs=as_core, inShellDefIdxs=in_shell)
    got, ref = np.zeros(4, np.float32), np.zeros(4, np.float32)
    dev.all_atoms_coord_number_coords(coordNumData=got, **kw)
    orc.all_atoms_coord_number_coords(coordNumData=ref, **kw)
    assert np.array_equal(got, ref) and ref.sum() > 1000
    idx = rng.integers(0, n, 13).astype(np.int32)                 # a molecule-sized group
    got, ref = np.ones(4, np.float32), np.ones(4, np.float32)
    dev.multi_atoms_coord_number_coords(indexes=idx, coordNumData=got, **kw)
    orc.multi_atoms_coord_number_coords(indexes=idx, coordNumData=ref, **kw)
    assert np.array_equal(got, ref)


@pytest.mark.gpu
def test_device_before_after_bookkeeping_equals_recount():
    """compute_before_move / compute_after_move of AtomicCoordinationNumberConstraint (AtomicCoordinationConstraints.py:519-577):
    the constraint halves the whole-system count (every core-shell pair is met from both of its ends, :505) and then
    data - before + after over a run of accepted moves equals the halved count from scratch of the final configuration"""
    from fullrmc_b200.Core import atomic_coordination as dev
    rng, box, basis, defs = _random_system(3000, 2, True, 3, 5)
    cores, shells, lowers, uppers, as_core, in_shell = defs
    kw = dict(basis=basis, isPBC=True, coresIndexes=cores, shellsIndexes=shells, lowerShells=lowers, upperShells=uppers,
              asCoreDefIdxs=as_core, inShellDefIdxs=in_shell)
    data = np.zeros(3, np.float32)
    dev.all_atoms_coord_number_coords(boxCoords=box, coordNumData=data, **kw)
    data /= np.float32(2.)
    for _ in range(12):
        idx = rng.integers(0, len(box), 1).astype(np.int32)      # single-atom moves: pairs inside a moved group would count once
        before, after = np.zeros(3, np.float32), np.zeros(3, np.float32)
        dev.multi_atoms_coord_number_coords(indexes=idx, boxCoords=box, coordNumData=before, **kw)
        box[idx] += (rng.normal(0, 0.02, (1, 3))).astype(np.float32)
        dev.multi_atoms_coord_number_coords(indexes=idx, boxCoords=box, coordNumData=after, **kw)
        data = data - before + after
    recount = np.zeros(3, np.float32)
    dev.all_atoms_coord_number_coords(boxCoords=box, coordNumData=recount, **kw)
    assert np.array_equal(data, recount / np.float32(2.)) and recount.sum() > 1000


@pytest.mark.gpu
def test_device_edge_cases():
    from fullrmc_b200.Core import atomic_coordination as dev
    box = np.array([[0.1, 0.1, 0.1], [0.2, 0.1, 0.1], [0.95, 0.1, 0.1]], np.float32)
    basis = (np.eye(3) * 10).astype(np.float32)
    empty = np.zeros(0, np.int32)
    assert dev.single_atom_single_shell_coords(0, empty, box, basis, True, 0.0, 5.0) == 0.0
    allidx = np.arange(3, dtype=np.int32)
    # the core atom itself is counted when lower <= 0 (the reference does not skip it); both shell ends are inclusive
    assert dev.single_atom_single_shell_coords(0, allidx, box, basis, True, 0.0, 5.0) == 3.0
    d01 = float(dev.single_atom_single_shell_totdists(np.array([0.5, 1.0, 1.5], np.float32), allidx, 1.0, 1.5))
    assert d01 == 2.0
    assert dev.single_atom_single_shell_coords(0, allidx, box, basis, False, 0.05, 0.2) == 1.0     # no images: only atom 1
    data = np.zeros(1, np.float32)
    dev.multi_atoms_coord_number_coords(indexes=empty, boxCoords=box, basis=basis, isPBC=True, coresIndexes=[allidx], shellsIndexes=[allidx],
                                        lowerShells=[0.5], upperShells=[2.0], asCoreDefIdxs=[[0]] * 3, inShellDefIdxs=[[0]] * 3, coordNumData=data)
    assert data[0] == 0.0
    with pytest.raises(ValueError):
        dev.single_atom_single_shell_coords(0, np.array([7], np.int32), box, basis, True, 0.0, 5.0)
    with pytest.raises(TypeError):
        dev.all_atoms_coord_number_totdists(np.zeros((3, 3), np.float32), [allidx], [allidx], [0.5], [2.0], [[0]] * 3, [[0]] * 3, data)


@pytest.mark.gpu
@pytest.mark.parametrize("pbc", [True, False])
def test_device_shell_bounds_that_are_distances_of_the_system(pbc, orc_pd):
    """both ends inclusive when the bounds ARE distances that occur (the device compares d^2 against thresholds found by
    exact fp32 search, csrc/coordnum.cu), and non-finite bounds fall back to the comparison on the rounded distance"""
    from fullrmc_b200.Core import atomic_coordination as dev
    from oracle import coordination as orc
    rng, box, basis, _ = _random_system(4000, 2, pbc, 2, 91)
    allidx = np.arange(len(box), dtype=np.int32)
    for core in (3, 1777):
        row = np.sort(orc_pd.pairs_distances_to_indexcoords(core, box, basis, pbc))
        for lo, up in ((row[40], row[400]), (row[0], row[1]), (row[7], row[7]), (np.float32(0), row[-1]), (row[100], row[99]),
                       (np.nextafter(row[40], np.float32(np.inf)), np.nextafter(row[400], np.float32(-np.inf))),
                       (np.float32(-1.0), np.float32(np.inf)), (np.float32(np.nan), row[50]), (row[5], np.float32(np.nan)),
                       (np.float32(-np.inf), row[9]), (np.float32(0), np.float32(0))):
            got = dev.single_atom_single_shell_coords(core, allidx, box, basis, pbc, lo, up)
            ref = orc.single_atom_single_shell_coords(core, allidx, box, basis, pbc, lo, up)
            assert got == ref, (core, lo, up, got, ref)
    assert orc.single_atom_single_shell_coords(3, allidx, box, basis, pbc, row[40], row[400]) >= 300


def test_host_flattening_lists_the_reference_visits():
    """CPU: the task table the mirror builds (vectorised, grouped by role) holds exactly the (atom, list, definition,
    bounds) visits of single_atom_coord_number_coords (:207-240) for every listed atom, lists deduplicated by identity"""
    from fullrmc_b200.Core import atomic_coordination as dev
    rng = np.random.default_rng(3)
    n, nT, ndef = 300, 3, 5
    el = rng.integers(0, nT, n).astype(np.int32)
    cores, shells, lowers, uppers, as_core, in_shell = definitions(rng, n, el, nT, ndef)
    shells[1] = cores[0]                                            # one array used twice: stored once
    in_shell = [[] for _ in range(n)]
    for d in range(ndef):
        for i in shells[d]:
            in_shell[int(i)].append(d)
    atoms = rng.integers(0, n, 40).tolist() + [-1]                  # a repeated and a negative index are legal
    lists = dev._Lists()
    core, lst, out, lo, up = dev._definition_tasks(atoms, atoms, cores, shells, lowers, uppers, as_core, in_shell, lists)
    off, idx = lists.flat()
    assert len(off) - 1 == 2 * ndef - 1 and off[-1] == idx.shape[0]
    got = sorted((int(a), tuple(idx[off[s]:off[s + 1]].tolist()), int(d), float(x), float(y)) for a, s, d, x, y in zip(core, lst, out, lo, up))
    want = []
    for a in atoms:
        for d in as_core[a]:
            want.append((a, tuple(shells[d].tolist()), d, float(lowers[d]), float(uppers[d])))
        for d in in_shell[a]:
            want.append((a, tuple(cores[d].tolist()), d, float(lowers[d]), float(uppers[d])))
    assert got == sorted(want) and len(got) > 40
    with pytest.raises(IndexError):
        dev._definition_tasks([0], [0], cores, shells, lowers, uppers, [[ndef]] * n, in_shell, dev._Lists())
    with pytest.raises(ValueError):
        dev._Lists().add(np.zeros(3, np.int64))                     # Cython would refuse the int64 buffer
