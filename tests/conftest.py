import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_available():
    try:
        from fullrmc_b200 import _lib
        return int(_lib.load_library().frmc_device_count()) > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a box without a CUDA device skips the gpu-marked tests instead of failing at the first one
    (the product has no CPU fallback: tests/test_abi.py::test_no_cpu_fallback_without_gpu)"""
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (fullrmc_b200 has no CPU fallback); run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_modules():
    """The reference's own compiled Cython modules (oracle/_ref), or None when not built.
    Built here from /root/reference by oracle/build_ref.py; travels to the GPU box as .so files."""
    from oracle import build_ref
    if not build_ref.is_built():
        build_ref.build()
    return build_ref.load()


@pytest.fixture(scope="session")
def orc():
    from oracle import pairhist
    pairhist.build()
    return pairhist


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
