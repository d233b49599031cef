"""One call, several GPUs: frmc_full_pairs_histograms_coords_multi (csrc/multigpu.cu) behind
fullrmc_b200.Core.pairs_histograms.full_pairs_histograms_coords(_devices=... / $FULLRMC_B200_DEVICES).

Runs on however many devices the box has (1: the single-device branch of the same entry point; the driver's GPU test
box and `gpurun --gpus 2` exercise the NCCL all-reduce).  Bars: identical to the one-device call and to the oracle."""
import numpy as np
import pytest

import cases as C

pytestmark = pytest.mark.gpu
CASES = {c["name"]: c for c in C.make_cases()}
HKEYS = ("basis", "isPBC", "moleculeIndex", "elementIndex", "numberOfElements", "minDistance", "maxDistance", "bin", "histSize")


def _devices():
    from fullrmc_b200 import _lib
    return list(range(min(8, int(_lib.load_library().frmc_device_count()))))


@pytest.mark.parametrize("name", ["ortho_atomic", "tri_molecular", "tri_unwrapped", "ibc_nanoparticle", "coincident_empty_class",
                                  "tiny_1", "cfg4_small"])
def test_multi_device_call_equals_oracle(name, orc):
    from fullrmc_b200 import _lib
    from fullrmc_b200.Core import pairs_histograms as ph
    case = CASES[name]
    kw = {k: case[k] for k in HKEYS}
    want = orc.full_pairs_histograms_coords(boxCoords=case["boxCoords"], ncores=orc.max_threads(), **kw)
    devices = _devices()
    for devs in ([devices[0]], devices, devices[::-1]):
        if len(devs) == 1:
            lib = _lib.load_library()                      # the multi entry point itself, one device
            import ctypes
            n, nEl, hs = case["boxCoords"].shape[0], case["numberOfElements"], case["histSize"]
            hi = np.empty((nEl, nEl, hs), np.float32); he = np.empty((nEl, nEl, hs), np.float32)
            ov = ctypes.c_uint64(0)
            dv = (ctypes.c_int * 1)(devs[0])
            rc = lib.frmc_full_pairs_histograms_coords_multi(1, dv, _lib.ptr(case["boxCoords"], _lib.c_f32p), n, _lib.ptr(case["basis"], _lib.c_f32p),
                                                             int(case["isPBC"]), _lib.ptr(case["moleculeIndex"], _lib.c_i32p),
                                                             _lib.ptr(case["elementIndex"], _lib.c_i32p), nEl, float(case["minDistance"]),
                                                             float(case["maxDistance"]), float(case["bin"]), hs, _lib.ptr(hi, _lib.c_f32p),
                                                             _lib.ptr(he, _lib.c_f32p), ctypes.byref(ov))
            _lib.check(rc, "multi(1)")
            got = (hi, he)
        else:
            got = ph.full_pairs_histograms_coords(boxCoords=case["boxCoords"], _devices=devs, **kw)
            assert _lib.load_library().frmc_multi_reduce_path().decode().startswith("nccl")
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), "devices %s" % devs


def test_multi_device_call_at_scale():
    """200 000 atoms (every device gets thousands of tasks): all devices together = one device, through the environment
    variable an unmodified Engine's process would carry"""
    import os
    from fullrmc_b200 import synthetic
    from fullrmc_b200.Core import pairs_histograms as ph
    s = synthetic.cfg5(200000)
    g = synthetic.RGrid(0.0, 0.02, 1000)
    kw = dict(s.hist_kwargs(), **g.kwargs())
    one = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    os.environ["FULLRMC_B200_DEVICES"] = "all"
    try:
        every = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    finally:
        del os.environ["FULLRMC_B200_DEVICES"]
    assert np.array_equal(one[0], every[0]) and np.array_equal(one[1], every[1])


def test_sliced_upload_on_five_or_more_devices():
    """from 5 devices on, the raw arrays reach device 0 in slices over every device's own PCIe link (csrc/multigpu.cu); a
    molecular system of 600 000 atoms (the molecule keys travel the same way): identical to the one-device call"""
    from fullrmc_b200 import synthetic
    from fullrmc_b200.Core import pairs_histograms as ph
    devices = _devices()
    if len(devices) < 5:
        pytest.skip("needs at least 5 devices")
    basis = np.array([[190, 0, 0], [12, 186, 0], [-9, 15, 188]], dtype=np.float32)
    s = synthetic.random_system(600000, 4, basis, n_elements=3, molecule_size=3)
    kw = dict(s.hist_kwargs(), minDistance=np.float32(0.0), maxDistance=np.float32(6.0), bin=np.float32(0.05), histSize=120)
    one = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, **kw)
    every = ph.full_pairs_histograms_coords(boxCoords=s.boxCoords, _devices=devices, **kw)
    assert np.array_equal(one[0], every[0]) and np.array_equal(one[1], every[1])
    assert one[0].sum() > 0 and one[1].sum() > 0
