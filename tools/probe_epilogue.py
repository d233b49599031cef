"""Debug probe: phase timing of the fused epilogue kernel (clock64 stamps) and ablations."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic, _lib as L
from fullrmc_b200.store import DeviceStore
from fullrmc_b200.model import ModelSpec

s = synthetic.cfg4()
grid = synthetic.RGrid(0.0, 0.02, 1000)
q = synthetic.q_values(nq=400)
common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume,
              rho0=s.numberDensity, shell_centers=grid.shellCenters, shell_volumes=grid.shellVolumes)
idx = np.array([5], np.int32)
for label, flag, kinds in (("normal", 1, ("PDF", "SQ")), ("skipSQ", 3, ("PDF", "SQ")), ("skipG", 5, ("PDF", "SQ")),
                           ("skipBoth", 7, ("PDF", "SQ")), ("noWait", 9, ("PDF", "SQ")), ("PDFonly", 1, ("PDF",)), ("SQonly", 1, ("SQ",)), ("nomodel", 1, ())):
    store = DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, 5)
    g = store.add_grid(grid.minDistance, grid.maxDistance, grid.bin, grid.hs)
    for k in kinds:
        if k == "PDF":
            store.add_model(g, ModelSpec("PDF", experimental=np.zeros(1000, np.float32), sq_exact=flag, **common))
        else:
            store.add_model(g, ModelSpec("SQ", experimental=np.ones(400, np.float32), q_values=q, sq_exact=flag, **common))
    store.compute_data()
    import ctypes as _ct
    from ctypes import c_void_p
    # reset timeline slots: min slot to a huge value
    store.propose(idx, s.boxCoords[idx] + np.float32(0.001)); store.reject()
    out = np.zeros(128, np.int64)
    L.check(store._lib.frmc_store_debug_stamps(store._handle, out.ctypes.data_as(L.c_i64p), 128), "stamps")
    ph = []
    for m in range(len(kinds)):
        st = out[m * 8:m * 8 + 6]; gt = out[64 + m * 8:64 + m * 8 + 6]
        ph.append([(int(st[i + 1] - st[i]), int(gt[i + 1] - gt[i])) if st[i + 1] and st[i] else None for i in range(4)])
    store.propose(idx, s.boxCoords[idx] + np.float32(0.001))
    ms = store.replay_proposal(300)
    tl = out[120:123]
    e0 = [int(out[64 + m * 8]) for m in range(len(kinds))]
    print("%-9s replay %.2f us/launch   phases %s" % (label, 1e3 * ms, ph))
    print("          timeline ns: delta start..end %d ; delta end -> epilogue stamp0 %s ; epilogue stamp0 -> last publish %s" % (
        int(tl[1] - tl[0]), [e - int(tl[1]) for e in e0], [int(tl[2]) - e for e in e0]))
    store.close()
