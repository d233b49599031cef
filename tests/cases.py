"""Seeded input cases shared by the golden-vector generator and the parity tests.

Each case is a dict of the arrays the reference extension functions take.  The cases
cover what the reference's callers produce (SURVEY.md section 8c "parity rules"):
orthorhombic / triclinic / non-periodic boxes, molecular and atomic systems, lattice
configurations with exact +-0.5 fractional differences, unwrapped fractional
coordinates (|frac| > 1: the Engine never wraps moved atoms, Engine.py:3223), coincident
atoms, empty element classes and tiny systems.
"""
import numpy as np

F32, I32 = np.float32, np.int32


def _case(name, box, basis, isPBC, mol, el, nEl, rmin, rmax, bin, hs):
    return dict(name=name, boxCoords=np.ascontiguousarray(box, dtype=F32), basis=np.ascontiguousarray(basis, dtype=F32),
                isPBC=bool(isPBC), moleculeIndex=np.ascontiguousarray(mol, dtype=I32),
                elementIndex=np.ascontiguousarray(el, dtype=I32), numberOfElements=int(nEl),
                minDistance=F32(rmin), maxDistance=F32(rmax), bin=F32(bin), histSize=int(hs))


def pdf_limits(r0, dr, hs):
    """PDF-style limits: min = r0 - dr/2, max = r_last + dr/2 (PairDistributionConstraints.py:747-748)."""
    r = (r0 + dr * np.arange(hs)).astype(F32)
    b = F32(r[1] - r[0])
    return F32(r[0] - b / 2.), F32(r[-1] + b / 2.), b


def make_cases():
    cases = []
    rng = np.random.default_rng(20261017)

    # 1. orthorhombic, wrapped, atomic (every atom its own molecule) -> ORTHO_FAST
    n = 1500
    box = rng.random((n, 3), dtype=F32)
    basis = np.diag([31.0, 29.5, 33.25]).astype(F32)
    rmin, rmax, b = pdf_limits(0.01, 0.02, 700)
    cases.append(_case("ortho_atomic", box, basis, True, np.arange(n), rng.integers(0, 3, n), 3, rmin, rmax, b, 700))

    # 2. triclinic, molecular (13-atom molecules, THF-like), wrapped -> TRI_FAST
    n = 1300
    box = rng.random((n, 3), dtype=F32)
    basis = np.array([[30, 0, 0], [4.5, 29, 0], [-3, 6, 28]], dtype=F32)
    cases.append(_case("tri_molecular", box, basis, True, np.arange(n) // 13, rng.integers(0, 3, n), 3,
                       0.5, 12.5, 0.02, 600))

    # 3. triclinic, unwrapped coordinates in [-1, 2) -> TRI_GEN
    n = 900
    box = (rng.random((n, 3), dtype=F32) * F32(3) - F32(1)).astype(F32)
    cases.append(_case("tri_unwrapped", box, basis, True, np.arange(n) // 5, rng.integers(0, 4, n), 4,
                       0.0, 13.0, 0.05, 260))

    # 4. orthorhombic unwrapped -> ORTHO_GEN
    n = 800
    box = (rng.random((n, 3), dtype=F32) * F32(4) - F32(1.5)).astype(F32)
    cases.append(_case("ortho_unwrapped", box, np.diag([25.0, 25.0, 25.0]).astype(F32), True, np.arange(n),
                       rng.integers(0, 2, n), 2, 1.0, 12.0, 0.1, 110))

    # 5. simple-cubic lattice: exact +-0.5 fractional differences and many distance ties on bin edges
    m = 8
    g = (np.arange(m, dtype=F32) / F32(m))
    box = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3).astype(F32)
    n = box.shape[0]
    cases.append(_case("lattice_half", box, np.diag([16.0, 16.0, 16.0]).astype(F32), True, np.arange(n) // 4,
                       np.arange(n) % 2, 2, 0.0, 8.0, 0.25, 32))

    # 6. non-periodic nanoparticle (SiOx-like), Cartesian coordinates, one molecule = everything -> IBC
    n = 1100
    pts = rng.normal(0.0, 9.0, (n, 3)).astype(F32)
    rmin, rmax, b = pdf_limits(0.02, 0.02, 1243 // 2)
    cases.append(_case("ibc_nanoparticle", pts, np.eye(3, dtype=F32), False, np.zeros(n), rng.integers(0, 2, n), 2,
                       rmin, rmax, b, 1243 // 2))

    # 7. coincident atoms + an element class with no atoms + duplicated positions across the cell
    n = 400
    box = rng.random((n, 3), dtype=F32)
    box[50:60] = box[40:50]                       # exact duplicates (distance 0 -> bin 0 when rmin = 0)
    box[100:110] = box[90:100] + F32(1.0)         # same site through one lattice translation
    el = rng.integers(0, 3, n) * 2 % 4            # elements {0,2}: classes 1 and 3 stay empty
    cases.append(_case("coincident_empty_class", box, np.diag([14.0, 15.0, 16.0]).astype(F32), True,
                       np.arange(n) // 2, el, 4, 0.0, 7.0, 0.05, 140))

    # 8. tiny systems (group-subset calls of compute_before_move use k = 1..13 atoms)
    for n in (1, 2, 13):
        box = rng.random((n, 3), dtype=F32)
        cases.append(_case("tiny_%d" % n, box, basis, True, np.zeros(n), rng.integers(0, 3, n), 3, 0.0, 14.0, 0.1, 140))

    # 9. larger than one I-tile of the R=4 kernel is exercised on the GPU box only (needs the oracle there);
    #    a 5-element cfg4-like box at reduced N keeps the golden file small
    n = 3000
    box = rng.random((n, 3), dtype=F32)
    basis4 = np.array([[31, 0, 0], [4.6, 30.4, 0], [-3.1, 6.2, 29.4]], dtype=F32)
    cases.append(_case("cfg4_small", box, basis4, True, np.arange(n), rng.integers(0, 5, n), 5, 0.0, 14.0, 0.02, 700))
    return cases


def group_for(case, rng):
    """A move group for the per-move path: one whole molecule when molecular, else one atom,
    plus (sometimes) a stranger from another molecule."""
    n = case["boxCoords"].shape[0]
    mol = case["moleculeIndex"]
    a = int(rng.integers(0, n))
    members = np.flatnonzero(mol == mol[a])
    if members.shape[0] > 16:
        members = members[:7]
    if rng.random() < 0.3 and n > members.shape[0] + 1:
        stranger = int(rng.integers(0, n))
        if stranger not in members:
            members = np.append(members, stranger)
    return members.astype(I32)
