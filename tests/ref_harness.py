"""Test scaffolding: import the UNMODIFIED reference package from /root/reference (constraint
classes + Engine) under the third-party stubs of tests/ref_stubs, with the reference's own
compiled kernels (oracle/_ref) in place of fullrmc.Core.<extension>.

Used by the golden-vector generators in the build container and by tests/test_dropin.py on the GPU box; nothing
of the reference is tracked by git: oracle/build_ref.stage_package copies the package and the example inputs into
the git-ignored oracle/_ref/, which travels to the GPU box like the built binaries.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF = os.environ.get("FULLRMC_REFERENCE", "/root/reference")


def examples_dir():
    """the shipped example inputs: /root/reference/Examples here, the staged copies on the GPU box"""
    from oracle import build_ref
    return os.path.join(REF, "Examples") if os.path.isdir(REF) else build_ref.EXAMPLES


DROPIN_MODULES = ("pairs_distances", "pairs_histograms", "reciprocal_space", "atomic_distances", "atomic_coordination")


def load_reference(dropin=False):
    """Import the unmodified reference package (staged by oracle/build_ref.stage_package under oracle/_ref/pkg, with
    the reference's own compiled kernels in fullrmc/Core) under the third-party stubs.

    dropin=True installs the CUDA backend the way SURVEY.md section 8b (ii) describes: ``sys.modules`` is pre-seeded
    with ``fullrmc_b200.Core.<name>`` under the names ``fullrmc.Core.<name>`` BEFORE anything imports
    ``fullrmc.Constraints.*``, so that the constraint modules' ``from ..Core.pairs_histograms import ...`` lines bind
    the CUDA functions.  No reference source is touched.  Must be the first import of fullrmc in the process."""
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    if not build_ref.stage_package():
        return None
    assert "fullrmc" not in sys.modules, "load_reference must run before anything imports fullrmc"
    sys.path.insert(0, build_ref.PKG)
    sys.path.insert(0, os.path.join(ROOT, "tests", "ref_stubs"))
    if dropin:
        import importlib
        for name in DROPIN_MODULES:
            sys.modules["fullrmc.Core." + name] = importlib.import_module("fullrmc_b200.Core." + name)
    import fullrmc  # noqa: F401
    if dropin:
        import fullrmc.Constraints.PairDistributionConstraints as pdm
        assert pdm.full_pairs_histograms_coords.__module__ == "fullrmc_b200.Core.pairs_histograms"
    return fullrmc


def fake_engine(fullrmc, boxCoordinates, basisVectors, isPBC, moleculesIndex, elementsIndex, elements):
    """An Engine instance whose private state is set directly: everything the
    three hot-path constraints read from their engine, nothing else."""
    from fullrmc.Engine import Engine
    from fullrmc.Core.Collection import Broadcaster, _AtomsCollector
    from fullrmc.Globals import FLOAT_TYPE
    E = object.__new__(Engine)
    n = boxCoordinates.shape[0]
    basis = np.ascontiguousarray(basisVectors, dtype=np.float32)
    box = np.ascontiguousarray(boxCoordinates, dtype=np.float32)
    if isPBC:
        real = (box.astype(np.float64) @ basis.astype(np.float64)).astype(np.float32)
        volume = FLOAT_TYPE(abs(np.linalg.det(basis.astype(np.float64))))
        rbasis = np.linalg.inv(basis.astype(np.float64)).astype(np.float32)
    else:
        real = box
        volume = FLOAT_TYPE(1. / 0.0333679 * n)
        rbasis = np.eye(3, dtype=np.float32)
    counts = np.bincount(elementsIndex, minlength=len(elements))
    allElements = [elements[i] for i in elementsIndex]
    priv = dict(repository=None, usedFrame="0", frames={"0": None}, constraints=[], broadcaster=Broadcaster(), state=1.0,
                boxCoordinates=box, realCoordinates=real, basisVectors=basis, reciprocalBasisVectors=rbasis,
                isPBC=bool(isPBC), isIBC=not bool(isPBC), moleculesIndex=np.ascontiguousarray(moleculesIndex, dtype=np.int32),
                elementsIndex=np.ascontiguousarray(elementsIndex, dtype=np.int32), allElements=allElements,
                elements=list(elements), numberOfAtomsPerElement={elements[i]: int(counts[i]) for i in range(len(elements))},
                volume=volume, numberDensity=FLOAT_TYPE(n) / FLOAT_TYPE(volume), accepted=0, generated=0, tried=0,
                numberOfAtoms=n, numberOfElements=len(elements),
                # what Engine.run / set_groups / set_group_selector touch beyond the constraints' needs (Engine.py:256-300)
                groups=[], groupSelector=None, tolerance=0., saveGroupsFlag=False, tolerated=0, removed=[0., 0., 0.],
                totalStandardError=None, lastSelectedGroupIndex=None, path=None, timeout=10, id="fake",
                pdb=type("PdbStandIn", (object,), {"numberOfAtoms": n})())     # add_group only asks it for numberOfAtoms
    for k, v in priv.items():
        object.__setattr__(E, "_Engine__" + k, v)
    object.__setattr__(E, "_runtime_ncores", np.int32(1))
    object.__setattr__(E, "_atomsCollector", _AtomsCollector(E))
    return E


def attach(E, constraint):
    constraint._set_engine(E)
    E._Engine__constraints.append(constraint)
    constraint.listen("engine set")
    return constraint
