"""Debug probe: device time per step of DeviceStore.run_generated on cfg4 / cfg5 (several calls in a row).
usage: python tools/probe_generated.py [cfg4|cfg5] [n_steps] [calls]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic
from fullrmc_b200.store import DeviceStore
from fullrmc_b200.model import ModelSpec

which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 4
F32 = np.float32
s = synthetic.cfg4() if which == "cfg4" else synthetic.cfg5()
grid = synthetic.RGrid(0.0, 0.02, 1000)
q = synthetic.q_values(nq=400)
common = dict(elements=s.elements, n_per_element=s.numberOfAtomsPerElement, weighting=s.weighting, volume=s.volume,
              rho0=s.numberDensity, shell_centers=grid.shellCenters, shell_volumes=grid.shellVolumes)
rng = np.random.default_rng(101)
smooth = lambda m, c: (c + 0.02 * np.convolve(rng.standard_normal(m + 20), np.ones(21) / 21.0, "valid")).astype(F32)
store = DeviceStore(s.boxCoords, s.basis, True, s.moleculeIndex, s.elementIndex, s.numberOfElements)
g = store.add_grid(grid.minDistance, grid.maxDistance, grid.bin, grid.hs)
store.add_model(g, ModelSpec("PDF", experimental=smooth(1000, 0.0), **common))
store.add_model(g, ModelSpec("SQ", experimental=smooth(400, 1.0), q_values=q, **common))
total = np.sum([F32(x) for x in store.compute_data()], dtype=F32)
b64 = s.basis.astype(np.float64)
store.set_groups(None)
store.set_real_coords((s.boxCoords.astype(np.float64) @ b64).astype(F32), np.linalg.inv(b64).astype(F32))
c = 0
for call in range(calls):
    l0 = store.batch_stats()
    t0 = time.perf_counter()
    out = store.run_generated(n, 7, c, 0.17, total)
    wall = time.perf_counter() - t0
    l1 = store.batch_stats()
    total = out["total"]; c += n
    print("%s call %d: %d steps, %d accepted, device %.2f us/step, wall %.2f us/step, %d launches, %d rounds" % (
        which, call, n, int((out["decisions"] > 0).sum()), 1e3 * out["device_ms"] / n, 1e6 * wall / n, l1[0] - l0[0], l1[1] - l0[1]))
