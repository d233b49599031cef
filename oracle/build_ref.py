"""Build the REAL reference kernels into oracle/_ref/ (test infrastructure, not product).

TEST INFRASTRUCTURE ONLY.  Nothing under ``fullrmc_b200/`` may import this.

The reference hot path is three Cython files that live, unmodified, under
``/root/reference/Extensions`` (pairs_distances.pyx, pairs_histograms.pyx,
reciprocal_space.pyx).  This script cythonizes them
*where they lie* (no reference source is copied into the repository: the
generated C goes to a temporary directory) and drops only the compiled
extension modules into ``oracle/_ref/fullrmc/Core/`` so that
``from fullrmc.Core.pairs_histograms import ...`` resolves exactly as it does
inside the reference package (pairs_histograms.pyx:11 imports
``fullrmc.Core.pairs_distances`` at module init).

Build recipe: gcc -O2, no -march, no -ffast-math,
-ffp-contract=off, Cython ``language_level=2`` and
``legacy_implicit_noexcept=True`` (the Cython 0.29 semantics the reference was
written for, README.md:62), no OpenMP (top-level setup.py:259-310 builds without
it, and the OpenMP path of pairs_histograms.pyx:53-68 is racy), ncores=1.

``oracle/_ref/`` is git-ignored but NOT gpurun-ignored, so the built ``.so``
files travel to the GPU box where /root/reference does not exist.
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FULLRMC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ("pairs_distances", "pairs_histograms", "reciprocal_space")
# only needed to IMPORT the reference's Engine/constraint classes (golden-vector generation); not on the path
EXTRA_MODULES = ("boundary_conditions_collection",)
# the distance-constraint kernels (atomic_distances.pyx imports pairs_distances)
# the coordination-number loops (atomic_coordination.pyx imports pairs_distances too)
NEXT_MODULES = ("atomic_distances", "atomic_coordination")


def is_built():
    core = os.path.join(OUT, "fullrmc", "Core")
    if not os.path.isdir(core):
        return False
    names = os.listdir(core)
    return all(any(n.startswith(m + ".") and n.endswith(".so") for n in names) for m in MODULES + NEXT_MODULES)


def build(force=False):
    """Compile the reference extensions.  Returns True when oracle/_ref is usable."""
    if is_built() and not force:
        return True
    ext_dir = os.path.join(REF, "Extensions")
    if not os.path.isdir(ext_dir):
        return False
    import numpy as np
    from setuptools import Distribution, Extension
    from Cython.Build import cythonize

    core = os.path.join(OUT, "fullrmc", "Core")
    os.makedirs(core, exist_ok=True)
    for d in (os.path.join(OUT, "fullrmc"), core):
        with open(os.path.join(d, "__init__.py"), "w") as fd:
            fd.write("# stub package so the compiled reference kernels import as fullrmc.Core.<name>\n")
    tmp = tempfile.mkdtemp(prefix="frmc_ref_build_")
    cwd = os.getcwd()
    old_cc = {k: os.environ.get(k) for k in ("CC", "LDSHARED")}
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["LDSHARED"] = "/usr/bin/gcc -shared"
    try:
        os.chdir(tmp)
        exts = [Extension("fullrmc.Core." + m,
                          [os.path.join(ext_dir, m + ".pyx")],
                          include_dirs=[np.get_include()],
                          extra_compile_args=["-O2", "-ffp-contract=off", "-w"],
                          define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])
                for m in MODULES + EXTRA_MODULES + NEXT_MODULES if os.path.exists(os.path.join(ext_dir, m + ".pyx"))]
        exts = cythonize(exts, build_dir=os.path.join(tmp, "cy"), language_level=2, quiet=True,
                         compiler_directives={"legacy_implicit_noexcept": True})
        dist = Distribution({"name": "fullrmc_ref_kernels", "ext_modules": exts})
        cmd = dist.get_command_obj("build_ext")
        cmd.build_lib = OUT
        cmd.build_temp = os.path.join(tmp, "obj")
        cmd.ensure_finalized()
        cmd.run()
    finally:
        os.chdir(cwd)
        for k, v in old_cc.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        shutil.rmtree(tmp, ignore_errors=True)
    return is_built()


PKG = os.path.join(OUT, "pkg")                     # the staged reference package: oracle/_ref/pkg/fullrmc
EXAMPLES = os.path.join(OUT, "Examples")          # the shipped inputs of BASELINE.json configs 1-3
EXAMPLE_FILES = {"atomicNiTi": ("system.pdb", "experimental.gr", "experimental.fq"),
                 "molecularTHF": ("thf.pdb", "thf_pdf.exp"),
                 "SiOxNanosphere": ("SiOx.pdb", "SiOx.gr")}
PY_DIRS = ("", "Core", "Constraints", "Generators", "Selectors")


def is_staged():
    return (os.path.exists(os.path.join(PKG, "fullrmc", "Engine.py")) and
            all(os.path.exists(os.path.join(EXAMPLES, d, f)) for d, fs in EXAMPLE_FILES.items() for f in fs))


def stage_package(force=False):
    """Stage the reference's pure-Python package (Engine, constraint classes, generators, selectors) next to its
    compiled kernels, plus the example inputs of configs 1-3, under the git-ignored oracle/_ref/ -- so that the
    UNMODIFIED reference classes and Engine.run can be driven on the GPU box, where /root/reference does not
    exist (tests/ref_harness.py; tests/test_dropin.py swaps fullrmc.Core.<extension> for fullrmc_b200.Core.<name>
    underneath them).  Binaries and inputs only travel; nothing here is tracked by git.  Returns True when usable."""
    if is_staged() and not force:
        return True
    if not os.path.isdir(REF) or not build():
        return False
    dst_pkg = os.path.join(PKG, "fullrmc")
    for sub in PY_DIRS:
        src_dir, dst_dir = os.path.join(REF, sub), os.path.join(dst_pkg, sub)
        os.makedirs(dst_dir, exist_ok=True)
        for name in os.listdir(src_dir):
            if name.endswith(".py") and name != "setup.py":
                shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
    so_dir = os.path.join(OUT, "fullrmc", "Core")
    for name in os.listdir(so_dir):
        if name.endswith(".so"):
            shutil.copyfile(os.path.join(so_dir, name), os.path.join(dst_pkg, "Core", name))
    for d, files in EXAMPLE_FILES.items():
        os.makedirs(os.path.join(EXAMPLES, d), exist_ok=True)
        for f in files:
            shutil.copyfile(os.path.join(REF, "Examples", d, f), os.path.join(EXAMPLES, d, f))
    return is_staged()


def load():
    """Import the compiled reference modules; returns (pairs_distances, pairs_histograms,
    reciprocal_space) or None when oracle/_ref has not been built."""
    if not is_built():
        return None
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import importlib
    return tuple(importlib.import_module("fullrmc.Core." + m) for m in MODULES)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built:", ok)
    staged = stage_package(force="--force" in sys.argv)
    print("reference package + example inputs staged under oracle/_ref:", staged)
    sys.exit(0 if ok else 1)
