"""The literal drop-in demonstration: the UNMODIFIED reference classes run on the CUDA backend.

`fullrmc.Core.pairs_histograms`, `pairs_distances`, `reciprocal_space`, `atomic_distances` and `atomic_coordination` are
replaced in `sys.modules` by the modules of `fullrmc_b200.Core` before the reference's constraint modules are imported
(tests/ref_harness.load_reference(dropin=True)); no line of the reference is changed.  Each test runs the fixture
GENERATOR itself that way in a fresh interpreter (the reference's own compiled kernels live in the pytest process), and
requires the files it writes to equal the committed fixtures -- generated in the build container with the reference's
own kernels -- array for array:

* tests/gen_golden_constraints.py  the reference constraint classes driven step by step (configs 1-3 + synthetic)
* tests/gen_golden_engine_run.py   whole Engine.run with fixed seeds (selector, generators, rigid pre-filter, Metropolis)

The reference package and the example inputs travel to the GPU box under the git-ignored oracle/_ref (oracle/build_ref.py).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run_generator(script, out_dir, names):
    from oracle import build_ref
    if not build_ref.stage_package():
        pytest.skip("reference package not staged (oracle/build_ref.py needs /root/reference once)")
    cmd = [sys.executable, os.path.join(ROOT, "tests", script), "--dropin", "--out", str(out_dir)] + list(names)
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert res.returncode == 0, res.stdout[-4000:]
    return res.stdout


def _same(a, b):
    if a.dtype.kind in "fc" and b.dtype.kind in "fc":
        return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
    return np.array_equal(a, b)


def _compare(path_new, path_golden):
    new, old = np.load(path_new), np.load(path_golden)
    assert sorted(new.files) == sorted(old.files)
    bad = [k for k in old.files if not _same(new[k], old[k])]
    assert not bad, "%s: arrays differ from the reference-generated fixture: %s" % (os.path.basename(path_golden), bad)


CLASS_CASES = ["niti", "thf", "siox", "synth", "niti_sf", "siox_shape", "synth_sf"]


def test_reference_constraint_classes_on_cuda_kernels(tmp_path, golden_dir):
    """PairDistribution / PairCorrelation / StructureFactor / ReducedStructureFactor constraints, unmodified, calling
    fullrmc_b200.Core.pairs_histograms: chi^2 of every move, scale-factor refits, shape-function refreshes, data arrays"""
    _run_generator("gen_golden_constraints.py", tmp_path, CLASS_CASES)
    for name in CLASS_CASES:
        _compare(os.path.join(str(tmp_path), "constraints_%s.npz" % name), os.path.join(golden_dir, "constraints_%s.npz" % name))


def test_reference_engine_run_on_cuda_kernels(tmp_path, golden_dir):
    """Engine.run with fixed seeds: the same accepted / tried counts, standard errors, scale factors, data and coordinates"""
    out = _run_generator("gen_golden_engine_run.py", tmp_path, ["niti", "thf", "siox"])
    for name in ("niti", "thf", "siox"):
        _compare(os.path.join(str(tmp_path), "engine_run_%s.npz" % name), os.path.join(golden_dir, "engine_run_%s.npz" % name))
    assert out.count("steps:") == 3
