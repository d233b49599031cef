"""Debug probe: one host round trip per Metropolis step (DeviceStore.step, acceptance on the host): wall time per step with
a launch per proposal and through the persistent kernel, cfg5 and cfg4.
usage: python tools/probe_step.py [n_steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullrmc_b200 import synthetic
from fullrmc_b200.store import DeviceStore
from fullrmc_b200.model import ModelSpec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
F32 = np.float32
for which in ("cfg5", "cfg4"):
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
    chi = store.compute_data()
    idx = rng.integers(0, s.numberOfAtoms, n + 200).astype(np.int32)
    inv = np.linalg.inv(s.basis.astype(np.float64))
    moved = (s.boxCoords[idx] + (rng.normal(0.0, 0.1, (n + 200, 3)) @ inv).astype(F32)).astype(F32)
    for mode in (False, True):
        store.set_persistent(mode)
        total = float(chi[0] + chi[1]); prev = None; acc = 0
        for it in range(n + 200):
            if it == 200:
                t0 = time.perf_counter()
            c = store.step(prev, idx[it:it + 1], moved[it:it + 1])
            t = float(c[0] + c[1])
            prev = t <= total
            if prev:
                total = t; acc += 1
        wall = time.perf_counter() - t0
        (store.accept if prev else store.reject)()
        print("%s persistent=%-5s %.2f us per step (%d accepted of %d)" % (which, mode, 1e6 * wall / n, acc, n + 200), flush=True)
    store.close()
