"""Test scaffolding: import the UNMODIFIED reference package from /root/reference (constraint
classes + Engine) under the third-party stubs of tests/ref_stubs, with the reference's own
compiled kernels (oracle/_ref) in place of fullrmc.Core.<extension>.

Used only by tests/gen_golden_constraints.py in the build container; nothing is copied into the
repository: the package tree is a directory of symlinks under a temporary directory.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FULLRMC_REFERENCE", "/root/reference")


def load_reference():
    if not os.path.isdir(REF):
        return None
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    assert build_ref.build(), "cannot build oracle/_ref"
    tmp = tempfile.mkdtemp(prefix="frmc_refpkg_")
    pkg = os.path.join(tmp, "fullrmc")
    os.makedirs(os.path.join(pkg, "Core"))
    for name in os.listdir(REF):
        if name != "Core":
            os.symlink(os.path.join(REF, name), os.path.join(pkg, name))
    for name in os.listdir(os.path.join(REF, "Core")):
        os.symlink(os.path.join(REF, "Core", name), os.path.join(pkg, "Core", name))
    so_dir = os.path.join(build_ref.OUT, "fullrmc", "Core")
    for name in os.listdir(so_dir):
        if name.endswith(".so"):
            os.symlink(os.path.join(so_dir, name), os.path.join(pkg, "Core", name))
    sys.path.insert(0, tmp)
    sys.path.insert(0, os.path.join(ROOT, "tests", "ref_stubs"))
    import fullrmc  # noqa: F401
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
                numberOfAtoms=n, numberOfElements=len(elements))
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
