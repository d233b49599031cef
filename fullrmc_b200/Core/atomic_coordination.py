"""Drop-in for ``fullrmc.Core.atomic_coordination`` (reference: Extensions/atomic_coordination.pyx), the counting
loops of AtomicCoordinationNumberConstraint (Constraints/AtomicCoordinationConstraints.py:492, :530, :561;
SURVEY.md section 8f rank 3).

Same function names, keyword names, argument meaning and in-place convention (``coordNumData`` is a float32 array
the counts are ADDED to).  Every function flattens its Python lists into one task table and makes ONE call to
``frmc_coordination_counts`` (csrc/coordnum.cu): the reference makes one Python-level call, a fancy-indexed copy
and a distance pass per (atom, definition).  Counts are integers, so the result is bit-identical to the reference
while a cell stays below 2**24 (beyond that the reference's float32 ``+= 1`` stops counting).  ``ncores`` is
accepted and ignored.

The reference's ``single_atom_coord_number_totdists`` (:171-198) walks ``asCoreDefIdxs`` in BOTH of its loops (the
coordinates form :207-240 walks ``inShellDefIdxs`` in the second); the mirror reproduces that as it is.
"""
from itertools import chain

import numpy as np

from .. import _lib as L

_F32, _I32, _I64 = np.float32, np.int32, np.int64


class _Lists(object):
    """atom lists gathered once per call: list object -> slot, offsets and one flat index array"""

    def __init__(self):
        self.slots, self.arrays, self.keep = {}, [], []

    def add(self, indexes):
        key = id(indexes)
        slot = self.slots.get(key)
        if slot is None:
            a = np.asarray(indexes)
            if a.ndim != 1:
                raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
            if a.size and a.dtype != _I32:
                raise ValueError("Buffer dtype mismatch, expected 'int32' but got '%s'" % a.dtype.name)
            slot = self.slots[key] = len(self.arrays)
            self.arrays.append(np.ascontiguousarray(a, dtype=_I32))
            self.keep.append(indexes)          # ids stay unique while the objects live
        return slot

    def flat(self):
        off = np.zeros(len(self.arrays) + 1, _I64)
        if self.arrays:
            off[1:] = np.cumsum([a.shape[0] for a in self.arrays])
            idx = np.concatenate(self.arrays) if off[-1] else np.zeros(0, _I32)
        else:
            idx = np.zeros(0, _I32)
        return off, np.ascontiguousarray(idx, dtype=_I32)


def _count(what, tasks, lists, nout, boxCoords=None, basis=None, isPBC=False, distances=None):
    """tasks: (core atom or distance row, list slot, output slot, lower, upper) as five equally long columns, or a
    list of such 5-tuples -> int32 counts [nout]"""
    lib = L.load_library()
    counts = np.zeros(nout, _I32)
    if isinstance(tasks, list):
        tasks = tuple(zip(*tasks)) if tasks else ((), (), (), (), ())
    core, lst, out = (np.ascontiguousarray(c, dtype=_I32) for c in tasks[:3])
    lower, upper = (np.ascontiguousarray(c, dtype=_F32) for c in tasks[3:])
    nt = core.shape[0]
    off, idx = lists.flat()
    if distances is None:
        coords = L.as_array(boxCoords, "boxCoords", _F32, 2)
        if coords.shape[1] != 3:
            raise ValueError("boxCoords must have 3 columns")
        b = np.ascontiguousarray(np.asarray(basis, dtype=_F32))
        if b.shape != (3, 3):
            raise ValueError("basis must be a (3,3) array")
        n, nrows, dist = coords.shape[0], 0, None
    else:
        dist = np.ascontiguousarray(distances, dtype=_F32)
        coords, b = None, None
        nrows, n = dist.shape
    rc = lib.frmc_coordination_counts(
        L.device_index(), L.ptr(coords, L.c_f32p), n, L.ptr(b, L.c_f32p), int(bool(isPBC)), L.ptr(dist, L.c_f32p), nrows,
        nt, L.ptr(core, L.c_i32p), L.ptr(lst, L.c_i32p), L.ptr(out, L.c_i32p), L.ptr(lower, L.c_f32p), L.ptr(upper, L.c_f32p),
        len(off) - 1, L.ptr(off, L.c_i64p), L.ptr(idx, L.c_i32p), nout, L.ptr(counts, L.c_i32p))
    L.check(rc, what)
    return counts


def _dist_row(distances, name="distances"):
    d = L.as_array(distances, name, _F32, 1)
    return d.reshape(1, -1)


# ---------------------------------------------------------------- one core atom
def single_atom_single_shell_subdists(distances, lowerShell, upperShell, ncores=1):
    """atomic_coordination.pyx:55-67 -- distances already restricted to the shell's atoms"""
    d = _dist_row(distances)
    lists = _Lists()
    slot = lists.add(np.arange(d.shape[1], dtype=_I32))
    return float(_count("single_atom_single_shell_subdists", [(0, slot, 0, lowerShell, upperShell)], lists, 1, distances=d)[0])


def single_atom_single_shell_totdists(distances, shellIndexes, lowerShell, upperShell, ncores=1):
    """atomic_coordination.pyx:71-85"""
    d = _dist_row(distances)
    lists = _Lists()
    slot = lists.add(shellIndexes)
    return float(_count("single_atom_single_shell_totdists", [(0, slot, 0, lowerShell, upperShell)], lists, 1, distances=d)[0])


def single_atom_single_shell_coords(coreIndex, shellIndexes, boxCoords, basis, isPBC, lowerShell, upperShell, ncores=1):
    """atomic_coordination.pyx:89-112"""
    lists = _Lists()
    slot = lists.add(shellIndexes)
    return float(_count("single_atom_single_shell_coords", [(int(coreIndex), slot, 0, lowerShell, upperShell)], lists, 1,
                        boxCoords=boxCoords, basis=basis, isPBC=isPBC)[0])


def single_atom_multi_shells_totdists(distances, shellsIndexes, lowerShells, upperShells, ncores=1):
    """atomic_coordination.pyx:116-135 -> float32 [len(shellsIndexes)]"""
    d = _dist_row(distances)
    lists = _Lists()
    tasks = [(0, lists.add(s), i, lowerShells[i], upperShells[i]) for i, s in enumerate(shellsIndexes)]
    return _count("single_atom_multi_shells_totdists", tasks, lists, len(shellsIndexes), distances=d).astype(_F32)


def single_atom_multi_shells_coords(coreIndex, shellsIndexes, boxCoords, basis, isPBC, lowerShells, upperShells, ncores=1):
    """atomic_coordination.pyx:139-167 -> float32 [len(shellsIndexes)]"""
    lists = _Lists()
    tasks = [(int(coreIndex), lists.add(s), i, lowerShells[i], upperShells[i]) for i, s in enumerate(shellsIndexes)]
    return _count("single_atom_multi_shells_coords", tasks, lists, len(shellsIndexes),
                  boxCoords=boxCoords, basis=basis, isPBC=isPBC).astype(_F32)


# ---------------------------------------------------------------- coordination-number definitions
def _definition_tasks(atoms, rows, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs, secondDefIdxs, lists):
    """the (atom, definition) visits of single_atom_coord_number_* for every atom of `atoms`, as task columns: an atom
    that is a core of a definition counts the definition's shell atoms, an atom named by `secondDefIdxs` counts its
    core atoms.  The order of the tasks is free (integer counts), so they are gathered per role, not per atom."""
    ndef = len(lowerShells)
    slot_shell = np.array([lists.add(shellsIndexes[d]) for d in range(ndef)], _I32)
    slot_core = np.array([lists.add(coresIndexes[d]) for d in range(ndef)], _I32)
    lower = np.array([lowerShells[d] for d in range(ndef)], _F32)
    upper = np.array([upperShells[d] for d in range(ndef)], _F32)
    rows = np.asarray(rows, _I32)
    core, lst, defs = [], [], []
    for defIdxs, slots in ((asCoreDefIdxs, slot_shell), (secondDefIdxs, slot_core)):
        per_atom = [defIdxs[a] for a in atoms]
        lens = np.fromiter(map(len, per_atom), _I64, len(per_atom))
        d = np.fromiter(chain.from_iterable(per_atom), _I32, int(lens.sum()))
        if d.size and (d.min() < 0 or d.max() >= ndef):
            raise IndexError("list index out of range")              # what lowerShells[defIdx] raises in the reference
        core.append(np.repeat(rows, lens))
        lst.append(slots[d])
        defs.append(d)
    d = np.concatenate(defs)
    return np.concatenate(core), np.concatenate(lst), d, lower[d], upper[d]


def _accumulate(coordNumData, counts):
    if not (isinstance(coordNumData, np.ndarray) and coordNumData.dtype == _F32 and coordNumData.ndim == 1):
        raise ValueError("Buffer dtype mismatch, expected 'float32' 1-d array for coordNumData")
    coordNumData += counts.astype(_F32)


def multi_atoms_coord_number_coords(indexes, boxCoords, basis, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells,
                                    asCoreDefIdxs, inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:280-313 (through :207-240): adds to coordNumData in place"""
    atoms = np.asarray(indexes).ravel().tolist()
    lists = _Lists()
    tasks = _definition_tasks(atoms, atoms, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs, inShellDefIdxs, lists)
    counts = _count("multi_atoms_coord_number_coords", tasks, lists, len(coordNumData), boxCoords=boxCoords, basis=basis, isPBC=isPBC)
    _accumulate(coordNumData, counts)


def single_atom_coord_number_coords(atomIndex, boxCoords, basis, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells,
                                    asCoreDefIdxs, inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:207-240"""
    multi_atoms_coord_number_coords(np.array([atomIndex], _I32), boxCoords, basis, isPBC, coresIndexes, shellsIndexes, lowerShells,
                                    upperShells, asCoreDefIdxs, inShellDefIdxs, coordNumData, ncores)


def all_atoms_coord_number_coords(boxCoords, basis, isPBC, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs,
                                  inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:349-376: every atom, as AtomicCoordinationNumberConstraint.compute_data calls it"""
    coords = L.as_array(boxCoords, "boxCoords", _F32, 2)
    multi_atoms_coord_number_coords(np.arange(coords.shape[0], dtype=_I32), coords, basis, isPBC, coresIndexes, shellsIndexes,
                                    lowerShells, upperShells, asCoreDefIdxs, inShellDefIdxs, coordNumData, ncores)


def multi_atoms_coord_number_totdists(indexes, distances, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs,
                                      inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:249-276 (through :171-198): distances[i] is the row of atom indexes[i]"""
    atoms = np.asarray(indexes).ravel().tolist()
    rows = [L.as_array(distances[i], "distances", _F32, 1) for i in range(len(atoms))]
    if not atoms:
        return
    d = np.ascontiguousarray(np.stack(rows))
    lists = _Lists()
    # the reference walks asCoreDefIdxs twice here (:185, :192); see the module docstring
    tasks = _definition_tasks(atoms, np.arange(len(atoms)), coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs, asCoreDefIdxs, lists)
    _accumulate(coordNumData, _count("multi_atoms_coord_number_totdists", tasks, lists, len(coordNumData), distances=d))


def single_atom_coord_number_totdists(atomIndex, distances, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs,
                                      inShellDefIdxs, coordNumData, ncores=1):
    """atomic_coordination.pyx:171-198"""
    multi_atoms_coord_number_totdists(np.array([atomIndex], _I32), [distances], coresIndexes, shellsIndexes, lowerShells,
                                      upperShells, asCoreDefIdxs, inShellDefIdxs, coordNumData, ncores)


def all_atoms_coord_number_totdists(distances, coresIndexes, shellsIndexes, lowerShells, upperShells, asCoreDefIdxs, inShellDefIdxs,
                                    coordNumData, ncores=1):
    """atomic_coordination.pyx:317-345 -- unusable in the reference: it passes its 2-d ndarray on to
    multi_atoms_coord_number_totdists, whose ``distances`` argument is typed ``list``, so every call ends in this
    TypeError (pinned by tests/golden/atomic_coordination.npz); kept for the name, with the same outcome"""
    raise TypeError("Argument 'distances' has incorrect type (expected list, got numpy.ndarray)")
